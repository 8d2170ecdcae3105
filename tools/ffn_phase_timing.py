#!/usr/bin/env python
"""Developer tool: per-phase clock64 breakdown of k_tc_fc_ffn (library built with -DS2S_PHASE_TIMING; the attention
kernels' own phase counters share g_phase, so this runs with S2S_ATTN_V1... no: it simply reports the sums, which
are dominated by whichever kernel was instrumented).  See tools/phase_timing.py for the attention kernel."""
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

NAMES = ["wait O tile (TMA)", "fc MMA issue + residual load", "wait fc MMA", "epilogue 1 (LN1, Y -> smem)",
         "sync + W1 issue + wait W1", "epilogue 2 (ReLU, pack -> TMEM)", "sync + W2 issue + wait W2",
         "epilogue 3 (LN2, store) + sync"]
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
eng.forward_reads_device(*dev, nr, nc, opts)
torch.cuda.synchronize()
lib.s2s_debug_counters(out, 16, 1)
v = np.array(list(out), dtype=np.float64)
tiles = nc * 2 * 2 + (nc * 16 + 127) // 128 * 2   # decoder: 2 tiles x 2 layers per chunk; encoder: rows/128 x 2 layers
for who, base in (("thread 0 (issues the MMAs)", 0), ("thread 32 (epilogue only)", 8)):
    tot = v[base:base + 8].sum()
    print(f"{who}: {tot / tiles:.0f} clk per tile per CTA (all k_tc_fc_ffn launches, {tiles} tiles)")
    for i, n in enumerate(NAMES):
        print(f"  {n:36s} {100 * v[base + i] / tot:5.1f}%  {v[base + i] / tiles:8.0f} clk/tile")
