#!/usr/bin/env python
"""Developer tool: in-situ per-kernel device time of one hot-path step (CUPTI through torch.profiler: launches are NOT
serialised and caches are NOT flushed, unlike the ncu launch list).  gpurun -- python tools/kernel_timeline.py"""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import default_opts, synth_reads  # noqa: E402
from seq2squiggle_b200.checkpoint import random_init_checkpoint, set_config  # noqa: E402
from seq2squiggle_b200.engine import Engine  # noqa: E402

cfg = set_config(None)
eng = Engine(random_init_checkpoint(cfg, 1)["state_dict"], cfg)
opts = default_opts("fp16")
b, ro, co = Engine.pack_reads(synth_reads(int(os.environ.get("READS", 4000)), seed=1), 9)
dev = [t.cuda() for t in (b, ro, co)]
nr, nc = ro.numel() - 1, int(co[-1])
for _ in range(3):
    eng.forward_reads_device(*dev, nr, nc, opts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e0.record()
    eng.forward_reads_device(*dev, nr, nc, opts)
    e1.record()
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and ev.device_time > 0:
        name = ev.name.split("(")[0].replace("s2s::", "").replace("(anonymous namespace)::", "")
        agg[name][0] += 1
        agg[name][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"# {nc} chunks, step {e0.elapsed_time(e1):.2f} ms (under CUPTI), kernel time {tot / 1e3:.2f} ms")
print(f"{'kernel':48s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'avg_us':>9s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:48]:48s} {n:8d} {t:10.1f} {100 * t / tot:6.1f}% {t / n:9.1f}")
