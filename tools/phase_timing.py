#!/usr/bin/env python
"""Developer tool: per-phase clock64 breakdown of k_tc_attn (library must be built with S2S_NVCC_EXTRA=-DS2S_PHASE_TIMING).
  S2S_NVCC_EXTRA=-DS2S_PHASE_TIMING python -c "from seq2squiggle_b200 import _lib; _lib.build(force=True)"
  gpurun -- python tools/phase_timing.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import default_opts, synth_reads  # noqa: E402
from seq2squiggle_b200 import _lib  # noqa: E402
from seq2squiggle_b200.checkpoint import random_init_checkpoint, set_config  # noqa: E402
from seq2squiggle_b200.engine import Engine  # noqa: E402

NAMES = ["0 wait X/W TMA", "1 wait QKV MMA", "2 QKV epilogue+sync", "3 wait S MMA", "4 max pass", "5 exp pass",
         "6 sync+wait PV MMA", "7 O epilogue+sync", "8", "9 prologue", "10 MMAw: wait P", "11 MMAw: issue PV",
         "12 MMAw: wait F", "13 MMAw: issue S", "14 MMAw: other"]

cfg = set_config(None)
eng = Engine(random_init_checkpoint(cfg, 1)["state_dict"], cfg)
lib = _lib.load()
opts = default_opts("fp16")
b, ro, co = Engine.pack_reads(synth_reads(int(os.environ.get("READS", 1000)), seed=1), 9)
dev = [t.cuda() for t in (b, ro, co)]
nr, nc = ro.numel() - 1, int(co[-1])
for _ in range(2):
    eng.forward_reads_device(*dev, nr, nc, opts)
out = (C.c_int64 * 16)()
lib.s2s_debug_counters(out, 16, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.forward_reads_device(*dev, nr, nc, opts)
e1.record()
torch.cuda.synchronize()
lib.s2s_debug_counters(out, 16, 1)
v = np.array(list(out), dtype=np.float64)
units = v[15]
tot = v[:10].sum()
NAMES = NAMES + [''] * 16
print(f"chunks {nc}, step {e0.elapsed_time(e1):.2f} ms, units(counted by tid0) {units:.0f}, clocks/unit {tot / max(units, 1):.0f}")
for i, n in enumerate(NAMES[:15]):
    if v[i]:
        per = v[i] / max(units, 1)
        print(f"  {n:24s} {100 * v[i] / tot:5.1f}%  {per:9.0f} clk/unit" + (f"  {per / 8:8.0f} clk/(head,tile)" if (3 <= i <= 7 or i >= 10) else ""))
