#!/usr/bin/env python
"""Developer tool: per-phase clock breakdown of k_tc_attn3 (library built with -DS2S_PHASE_TIMING=3).
  S2S_LIB_PATH=$PWD/seq2squiggle_b200/libs2s_b200_phase.so S2S_NVCC_EXTRA=-DS2S_PHASE_TIMING=3 \
      python -c "from seq2squiggle_b200 import _lib; _lib.build(force=True)"
  gpurun -- env S2S_LIB_PATH=... python tools/attn3_phase_timing.py"""
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

NAMES = {0: "softmax: wait S (+ld issue)", 1: "softmax: tcgen05.wait::ld", 2: "softmax: exp + st issue",
         3: "softmax: (unused)", 4: "softmax: wait::st + arrive P", 5: "softmax: unit boundary",
         6: "MMA warp: wait P", 7: "MMA warp: issue P.V", 8: "MMA warp: wait P.V done", 9: "MMA warp: issue S",
         10: "MMA warp: wait O read", 11: "MMA warp: wait K/V (unit start)", 12: "output warp: wait P.V of the last quarter", 13: "output warp: read / normalise / store", 14: "producer: wait QKV MMA",
         15: "producer: epilogue"}

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
units = 2.0 * nc * 2   # (chunk, head group) x 2 decoder layers
print(f"chunks {nc}, step {e0.elapsed_time(e1):.2f} ms, units {units:.0f}")
for lo, hi, who in ((0, 6, "softmax warp 0"), (6, 12, "MMA warp (tile 0)"), (12, 14, "output warp"), (14, 16, "producer warp")):
    tot = v[lo:hi].sum()
    print(f"{who}: {tot / units:.0f} clk/unit")
    for i in range(lo, hi):
        print(f"  {NAMES.get(i, str(i)):40s} {100 * v[i] / max(tot, 1):5.1f}%  {v[i] / units:8.0f} clk/unit  {v[i] / units / 16:7.0f} clk/quarter")
