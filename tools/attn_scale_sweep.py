#!/usr/bin/env python
"""Developer tool: robustness of the pipelined attention against sharp attention.  W_q / W_k of the decoder layers of
the random-init checkpoint are multiplied by `scale` (scores grow with scale^2); reports throughput and the fraction of
(chunk, head group) units that the fast kernel flagged and the exact kernel recomputed.
  gpurun -- python tools/attn_scale_sweep.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import default_opts, synth_reads  # noqa: E402
from seq2squiggle_b200 import _lib  # noqa: E402
from seq2squiggle_b200.checkpoint import random_init_checkpoint, set_config  # noqa: E402
from seq2squiggle_b200.engine import Engine  # noqa: E402

cfg = set_config(None)
base = random_init_checkpoint(cfg, 1)["state_dict"]
lib = _lib.load()
opts = default_opts("fp16")
b, ro, co = Engine.pack_reads(synth_reads(int(os.environ.get("READS", 2000)), seed=1), 9)
dev = [t.cuda() for t in (b, ro, co)]
nr, nc = ro.numel() - 1, int(co[-1])
counters = (C.c_int64 * 16)()
print(f"k_tc_attn3 (reference = the row's own score); {nc} chunks")
for scale in (1.0, 2.0, 3.0, 4.0, 6.0, 9.0):
    sd = dict(base)
    for layer in range(cfg["decoder_layers"]):
        for name in ("w_qs", "w_ks"):
            for part in ("weight", "bias"):
                key = f"decoders.layer_stack_FFT.{layer}.slf_attn.{name}.{part}"
                sd[key] = sd[key] * scale
    eng = Engine(sd, cfg)
    for _ in range(2):
        eng.forward_reads_device(*dev, nr, nc, opts)
    torch.cuda.synchronize()
    lib.s2s_debug_counters(counters, 16, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.forward_reads_device(*dev, nr, nc, opts)
    e1.record()
    torch.cuda.synchronize()
    eng.check()
    lib.s2s_debug_counters(counters, 16, 1)
    ms = e0.elapsed_time(e1)
    units = 2 * nc * cfg["decoder_layers"]
    print(f"scale {scale:4.1f}: {nc / ms / 1e3:6.3f} M chunks/s, flagging warps {counters[12]:8d} of {4 * units} ({100 * counters[12] / (4 * units):5.1f} %)")
    eng.close()
