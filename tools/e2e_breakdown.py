#!/usr/bin/env python
"""Developer tool: where the end-to-end (host reads -> host int16) time of model.predict_reads goes.  Warm-up of W full
steps, then REPS repetitions of STEPS steps, each on a fresh pipeline (like bench.py); with S2S_PIPE_TRACE=1 the slowest
repetition's per-piece trace (host timestamps + device events) is printed.
  S2S_PIPE_TRACE=1 gpurun -- python tools/e2e_breakdown.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_reads  # noqa: E402
from seq2squiggle_b200 import model as M  # noqa: E402
from seq2squiggle_b200.checkpoint import random_init_checkpoint, set_config  # noqa: E402
from seq2squiggle_b200.profiles import get_profile  # noqa: E402


class Sink:
    profile, profile_name = get_profile("dna-r10-prom"), "dna-r10-prom"
    signals, samples = None, 0

    def save(self):
        self.samples += sum(len(v) for v in self.signals.values())


cfg = set_config(None)
sd = random_init_checkpoint(cfg, 1)["state_dict"]
sink = Sink()
m = M.seq2squiggle(config=cfg, state_dict=sd, out_writer=sink, dwell_mean=12.5, dwell_std=0.0, noise_std=2.0,
                   noise_sampling=True, duration_sampling=True, min_noise=0.0, min_duration=3, device=0, seed=7)
steps, reps, warm = int(os.environ.get("STEPS", 5)), int(os.environ.get("REPS", 6)), int(os.environ.get("WARMUP", 3))
host_reads = [[(r.decode("latin-1"), str(j)) for j, r in enumerate(synth_reads(4000, seed=i))] for i in range(steps + warm)]
for rd in host_reads[:warm]:
    m.predict_reads(rd)
m.on_predict_epoch_end()
worst = None
for rep in range(reps):
    sink.samples = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    per = []
    for rd in host_reads[warm:]:
        ta = time.perf_counter()
        m.predict_reads(rd)
        per.append(1e3 * (time.perf_counter() - ta))
    pipe = m._pipe
    t1 = time.perf_counter()
    m.on_predict_epoch_end()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"rep {rep}: per-step submit ms {' '.join(f'{x:.0f}' for x in per)}; drain {1e3 * (t2 - t1):.0f} ms; "
          f"{sink.samples / (t2 - t0) / 1e6:.1f} M samples/s; staging allocs {pipe.stats['allocs']}")
    if pipe.trace and (worst is None or t2 - t0 > worst[0]):
        worst = (t2 - t0, rep, t0, list(pipe.trace))
    pipe.trace.clear()
if worst:
    _, rep, t0, tr = worst
    e0 = tr[0]["ev_start"]
    print(f"slowest repetition: {rep}")
    print("piece:  host submit@ms  pack ms  enqueue ms | fetch wait ms | device start@ms  device ms  idle-before ms")
    prev_end = 0.0
    for i, x in enumerate(tr):
        ds, de = e0.elapsed_time(x["ev_start"]), e0.elapsed_time(x["ev_done"])
        print(f"{i:3d}  {1e3 * (x['t_sub'] - t0):9.1f} {1e3 * (x['t_packed'] - x['t_sub']):8.1f} {1e3 * (x['t_enq'] - x['t_packed']):8.1f} | "
              f"{1e3 * (x['t_fetch1'] - x['t_fetch0']):8.1f} | {ds:9.1f} {de - ds:8.1f} {ds - prev_end:8.1f}")
        prev_end = de
