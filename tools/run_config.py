#!/usr/bin/env python
"""Developer tool: run one BASELINE.json configuration end to end through ``inference_run`` (read sampling, hot path,
BLOW5 writer) on the GPU box and print where the wall-clock time goes.  The lambda genome is not available on the
box, so a synthetic genome of the same length (or ``--genome-len``) is written first.

  gpurun -- python tools/run_config.py --config 1 --n 100000
  gpurun -- python tools/run_config.py --config 2 --n 20000        (read mode, 10 ragged reads sampled n times)
  gpurun -- python tools/run_config.py --config 3 --n 100000       (dna-r9-min, k = 6)
  gpurun -- python tools/run_config.py --config 4 --genome-len 100000000 --c 1 --r 5000
  gpurun --gpus 2 -- python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
         tools/run_config.py --config 4 --genome-len 100000000 --coverage 3 --read-length 5000 --out /tmp/sim.blow5
         (read-sharded; torchrun's own parser claims abbreviations such as --r, hence the long spellings)
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seq2squiggle_b200 import inference, reads as reads_mod  # noqa: E402
from seq2squiggle_b200.checkpoint import random_init_checkpoint, set_config  # noqa: E402
from seq2squiggle_b200.cli import set_seeds  # noqa: E402
from seq2squiggle_b200.profiles import update_config  # noqa: E402

T = {}


def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        return r
    return w


def blow5_summary(path):
    """(records, samples) of an uncompressed BLOW5 file, streaming (record layout: tests/blow5_reader.py)."""
    import struct
    nrec = nsamp = 0
    with open(path, "rb") as f:
        head = f.read(64)
        assert head[:6] == b"BLOW5\x01" and head[9] == 0
        (hsize,) = struct.unpack("<I", f.read(4))
        f.seek(hsize, 1)
        while True:
            szb = f.read(8)
            if len(szb) < 8 or szb[:5] == b"5WOLB":
                break
            (size,) = struct.unpack("<Q", szb)
            (idl,) = struct.unpack("<H", f.read(2))
            f.seek(idl + 4 + 32, 1)
            (n,) = struct.unpack("<Q", f.read(8))
            f.seek(size - (2 + idl + 4 + 32 + 8), 1)
            nrec += 1
            nsamp += n
    return nrec, nsamp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--c", "--coverage", dest="c", type=int, default=-1)
    ap.add_argument("--r", "--read-length", dest="r", type=int, default=1000)
    ap.add_argument("--genome-len", type=int, default=48502)
    ap.add_argument("--fasta", default=None, help="use this genome / reads file (e.g. tests/golden/lamda_genome.fasta) "
                    "instead of writing a synthetic one")
    ap.add_argument("--out", default=None, help="output file (default: a temporary .blow5; /dev/null is not seekable)")
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--profile", action="store_true", help="cProfile the main thread of inference_run")
    ap.add_argument("--null-sink", action="store_true",
                    help="writer -> nowhere: records are encoded (and counted) but not written (the BASELINE configs[4] 'null sink' figure)")
    a = ap.parse_args()
    profile = "dna-r9-min" if a.config == 3 else "dna-r10-prom"
    cfg = update_config(profile, set_config(None))
    tmp = tempfile.mkdtemp(prefix="s2s_cfg_")
    ckpt = os.path.join(tmp, "random_init.ckpt")
    import torch
    torch.save(random_init_checkpoint(cfg, seed=1), ckpt)
    rng = np.random.default_rng(3)
    fasta = os.path.join(tmp, "input.fasta")
    read_input = a.config == 2
    if a.fasta:
        fasta = a.fasta
    with open(os.devnull if a.fasta else fasta, "w") as f:
        if a.fasta:
            pass
        elif read_input:   # the 10 read lengths of example/lamda_genome_reads.fasta
            for i, ln in enumerate([2843, 10510, 8487, 2207, 11936, 4407, 1434, 2449, 14971, 11072]):
                f.write(f">read{i}\n" + rng.choice(list("ACGT"), ln).astype("U1").tobytes().decode("utf-32-le") + "\n")
        else:
            g = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), a.genome_len).tobytes().decode()
            f.write(">chr1 synthetic\n")
            for i in range(0, len(g), 70):
                f.write(g[i:i + 70] + "\n")
    out = a.out or os.path.join(tmp, "sim.blow5")
    inference.get_reads = timed("get_reads (genome preprocessing + read sampling)", reads_mod.get_reads)
    inference.get_reads_batches = timed("get_reads_batches (genome preprocessing + lengths-only replay + batch plan)", reads_mod.get_reads_batches)
    from seq2squiggle_b200 import model as M
    M.seq2squiggle.predict_reads = timed("predict_reads (pack + enqueue, blocks on the previous batch)", M.seq2squiggle.predict_reads)
    M.seq2squiggle.on_predict_epoch_end = timed("on_predict_epoch_end (drain + writer join)", M.seq2squiggle.on_predict_epoch_end)
    M.seq2squiggle.load_from_checkpoint = classmethod(timed("load_from_checkpoint (incl. CUDA context)", M.seq2squiggle.load_from_checkpoint.__func__))
    if a.null_sink:      # keep every byte of host work but the pwrite / fwrite itself
        os.environ["S2S_BLOW5_NULL_SINK"] = "1"
    set_seeds(a.seed)
    prof = None
    if a.profile:
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    t0 = time.perf_counter()
    import torch.cuda as _tc
    _orig_set_device = _tc.set_device
    _tc.set_device = timed("torch.cuda.set_device (CUDA initialisation)", _orig_set_device)
    inference.inference_run(config=cfg, saved_weights=ckpt, fasta=fasta, read_input=read_input, n=a.n if a.c < 0 else -1,
                            r=a.r, c=a.c, out=out, profile=profile, dwell_mean=None, dwell_std=0.0, noise_std=2.0,
                            noise_sampling=True, duration_sampling=True, distr="expon", predict_batch_size=1024,
                            export_every_n_samples=2000000, sample_rate=None, bps=None, digitisation=None, range_val=None,
                            offset_mean=None, offset_std=None, median_before_mean=None, median_before_std=None,
                            min_noise=0.0, min_duration=3, min_read_len=30, preserve_read_ids=False, seed=a.seed)
    wall = time.perf_counter() - t0
    if int(os.environ.get("RANK", "0")) != 0:      # sharded run: all ranks wrote into the one file; rank 0 reports
        print(f"rank {os.environ['RANK']}: {wall:.2f} s wall; " + "; ".join(f"{k.split(' (')[0]} {v:.2f} s" for k, v in T.items()))
        return
    if prof:
        import pstats
        prof.disable()
        pstats.Stats(prof).sort_stats("tottime").print_stats(18)
    size = os.path.getsize(out) if os.path.exists(out) else 0
    if a.null_sink:
        print(f"null sink: {wall:.2f} s wall; " + "; ".join(f"{k.split(' (')[0]} {v:.2f} s" for k, v in T.items()))
        return
    nrec, nsamp = blow5_summary(out)
    print(f"config {a.config} ({profile}, {'read' if read_input else 'reference'} mode): {nrec} reads, {nsamp} samples, "
          f"{size / 1e6:.1f} MB BLOW5 in {wall:.2f} s wall = {nrec / wall:.0f} reads/s, {nsamp / wall / 1e6:.1f} M samples/s "
          f"(whole command incl. model load and read sampling)")
    for k, v in T.items():
        print(f"  {k:62s} {v:8.2f} s")
    hot = T.get("predict_reads (pack + enqueue, blocks on the previous batch)", 0) + T.get("on_predict_epoch_end (drain + writer join)", 0)
    if hot:
        print(f"  predict loop only: {nsamp / hot / 1e6:.1f} M samples/s, {nrec / hot:.0f} reads/s")


if __name__ == "__main__":
    main()
