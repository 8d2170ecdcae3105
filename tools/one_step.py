#!/usr/bin/env python
"""Developer tool: STEPS passes of the hot path over READS synthetic reads (device-resident inputs), for ncu.
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/one_step.py"""
import os
import sys

import torch

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
for _ in range(int(os.environ.get("STEPS", 1))):
    raw, off, _ = eng.forward_reads_device(*dev, nr, nc, opts)
torch.cuda.synchronize()
eng.check()
print(f"{nc} chunks, {int(off[-1])} samples")
