#!/usr/bin/env python
"""Developer tool: where the end-to-end (host reads -> host int16) time of model.predict_reads goes.
  gpurun -- python tools/e2e_breakdown.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_reads  # noqa: E402
from seq2squiggle_b200 import model as M  # noqa: E402
from seq2squiggle_b200.checkpoint import random_init_checkpoint, set_config  # noqa: E402
from seq2squiggle_b200.engine import Engine  # noqa: E402
from seq2squiggle_b200.profiles import get_profile  # noqa: E402


class Sink:
    profile, profile_name = get_profile("dna-r10-prom"), "dna-r10-prom"
    signals, samples, t = None, 0, 0.0

    def save(self):
        t0 = time.perf_counter()
        self.samples += sum(len(v) for v in self.signals.values())
        self.t += time.perf_counter() - t0


T = {}


def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        return r
    return w


cfg = set_config(None)
sd = random_init_checkpoint(cfg, 1)["state_dict"]
sink = Sink()
m = M.seq2squiggle(config=cfg, state_dict=sd, out_writer=sink, dwell_mean=12.5, dwell_std=0.0, noise_std=2.0,
                   noise_sampling=True, duration_sampling=True, min_noise=0.0, min_duration=3, device=0, seed=7)
steps = int(os.environ.get("STEPS", 8))
host_reads = [[(r.decode("latin-1"), str(j)) for j, r in enumerate(synth_reads(4000, seed=i))] for i in range(steps)]
m.predict_reads(host_reads[0][:64])
m.on_predict_epoch_end()
Engine.pack_reads = staticmethod(timed("pack_reads", Engine.pack_reads))
M._ReadPipeline._fetch = timed("fetch(prev): wait offsets + queue D2H", M._ReadPipeline._fetch)
M._ReadPipeline.submit = timed("submit total", M._ReadPipeline.submit)
m.engine.forward_reads_device = timed("forward_reads_device (enqueue)", m.engine.forward_reads_device)
sink.samples = 0
torch.cuda.synchronize()
t0 = time.perf_counter()
per = []
for rd in host_reads:
    ta = time.perf_counter()
    m.predict_reads(rd)
    per.append(1e3 * (time.perf_counter() - ta))
print("per-step submit ms:", " ".join(f"{x:.0f}" for x in per))
t1 = time.perf_counter()
m.on_predict_epoch_end()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{steps} steps: submit loop {1e3 * (t1 - t0):.1f} ms, drain {1e3 * (t2 - t1):.1f} ms, total/step {1e3 * (t2 - t0) / steps:.1f} ms, "
      f"{sink.samples / (t2 - t0) / 1e6:.1f} M samples/s; sink.save {1e3 * sink.t:.1f} ms")
for k, v in T.items():
    print(f"  {k:42s} {1e3 * v / steps:8.2f} ms/step")
