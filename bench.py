#!/usr/bin/env python
"""Benchmark of the seq2squiggle predict hot path on B200 (contract: see the task description / DESIGN.md).

One "step" = one pass of the whole hot path (tokenise -> embed -> encoder -> samplers -> length regulator ->
decoder -> noise -> zero-strip -> digitise -> per-read compaction) over one batch of synthetic reads drawn from
the reference's read-length distribution (utils.py:325-331, expon, -r 1000) on a synthetic 48,502-bp genome
(the lambda genome's length): BASELINE.json configs[1] "lambda genome, -n 100000, default noise + duration
samplers", random-init checkpoint of the default architecture.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads-per-step R]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, reads sharded, no collective on the path)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GENOME_LEN = 48502
# algorithmic work per chunk: MMA FLOPs 2*M*N*K at k = 9
FLOP_PER_CHUNK = 85_083_392
ATT_FLOP_PER_CHUNK_LAYER = 16_000_000       # QK^T 8.0 M + PV 8.0 M per decoder layer
ATT_EXP_PER_CHUNK_LAYER = 8 * 250 * 250     # softmax exponentials per decoder layer
MUFU_PER_CLK_SM = 16                        # ex2 per clock per SM (4 per SM sub-partition)
# k_tc_attn2, one launch of 32768 chunks, `ncu --set full` (profiles/r01_attn2_ncu.txt): dram__bytes_read.sum 1.090957 GB
# + dram__bytes_write.sum 1.043353 GB; the algorithmic traffic is 2 x 256 rows x 128 B = 65,536 B per chunk (x16 in, o16 out)
ATT_DRAM_BYTES_PER_CHUNK_NCU = (1.090957e9 + 1.043353e9) / 32768
ATT_ALGO_BYTES_PER_CHUNK = 2 * 256 * 128


def synth_reads(n_reads: int, seed: int, r: int = 1000):
    """Reads with the reference's 'expon' length law (utils.py:325-331) from a synthetic lambda-sized genome."""
    rng = np.random.default_rng(seed)
    genome = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=GENOME_LEN)
    out = []
    while len(out) < n_reads:
        n = n_reads - len(out)
        ln = (213.98910256668592 + rng.exponential(6972.5319847131141, size=2 * n)) * r / 7106.0
        ln = np.clip(ln.astype(np.int64), 1, GENOME_LEN)
        st = rng.integers(0, GENOME_LEN, size=2 * n)
        ok = (st + ln <= GENOME_LEN) & (ln >= 30)             # read_check: full length inside the genome, >= 30 nt
        for s, l in zip(st[ok][:n], ln[ok][:n]):
            out.append(genome[s:s + l].tobytes())
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def default_opts(precision: str, seed: int = 7):
    from seq2squiggle_b200.engine import RunOptions
    from seq2squiggle_b200.profiles import get_profile
    # CLI defaults of `seq2squiggle predict` (seq2squiggle.py:230-390): samplers on, noise-std 2.0, min_duration 3
    return RunOptions.from_profile(get_profile("dna-r10-prom"), "dna-r10-prom", duration_sampling=True, dwell_std=0.0,
                                   noise_std=2.0, noise_sampling=True, min_noise=0.0, min_duration=3, seed=seed,
                                   precision=precision)


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_step(sd, cfg, reads, batch_chunks=1024):
    """One bounded sample of the reference CPU path: tokenise (utils.py:350-356, the reference's Python loops),
    predict_step per DataLoader batch of 1024 chunks (model.py:195-250), export + digitise.  Returns
    (emitted samples, chunks, seconds)."""
    import torch
    from oracle import s2s_oracle as orc
    from oracle.profiles_kat import PROFILES
    prof = PROFILES["dna-r10-prom"]
    t0 = time.perf_counter()
    ids, chunks = [], []
    for i, r in enumerate(reads):
        c = orc.split_sequence(r.decode("latin-1"), cfg)
        if c.size:
            chunks.append(c)
            ids += [i] * len(c)
    data = torch.from_numpy(np.concatenate(chunks, 0))
    preds = []
    with torch.inference_mode():
        for b in range(0, data.shape[0], batch_chunks):
            preds.append(orc.predict_step(sd, cfg, data[b:b + batch_chunks], dwell_mean=12.5, dwell_std=0.0,
                                          noise_std=2.0, noise_sampling=True, duration_sampling=True, min_noise=0.0,
                                          min_duration=3))
    sig = orc.assemble_reads(ids, torch.cat(preds))
    n = 0
    for s in sig.values():
        n += len(orc.digitise(s.reshape(-1).numpy(), prof["digitisation"], prof["range"], prof["offset_mean"]))
    return n, data.shape[0], time.perf_counter() - t0


def run_reference(args):
    import torch
    from seq2squiggle_b200.checkpoint import random_init_checkpoint, set_config
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = set_config(None)
    sd = random_init_checkpoint(cfg, seed=1)["state_dict"]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_reads = args.ref_reads
    batches = [synth_reads(n_reads, seed=100 + i) for i in range(args.warmup + args.steps)]
    for i in range(args.warmup):
        cpu_reference_step(sd, cfg, batches[i])
    tot_s, tot_c, tot_t, tot_r = 0, 0, 0.0, 0
    for i in range(args.warmup, args.warmup + args.steps):
        n, c, t = cpu_reference_step(sd, cfg, batches[i])
        tot_s += n; tot_c += c; tot_t += t; tot_r += n_reads
    value = tot_s / tot_t
    sample = f"{args.steps} steps x {n_reads} reads (~{tot_c // max(args.steps, 1)} chunks/step) of the same read distribution"
    line = {"impl": "reference", "metric": "simulated raw-signal samples/sec", "value": value, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n_reads, None),
            "reads_per_s": tot_r / tot_t, "chunks_per_s": tot_c / tot_t,
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_result(line)


def workload_config(reads_per_step, batch_chunks):
    return {"workload": "configs[1]: lambda-sized genome (48,502 bp, synthetic ACGT), reference mode, expon read "
                        "lengths -r 1000, dna-r10-prom, duration+noise samplers on, noise-std 2.0, random-init "
                        "default architecture (k=9, d=64, 2+2 FFT blocks)",
            "reads_per_step_per_gpu": reads_per_step, "sub_batch_chunks": batch_chunks,
            "l2": "per-step working set (>1 GB of activations) exceeds the 126 MB L2; a different read batch every step"}


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from seq2squiggle_b200 import _lib
    from seq2squiggle_b200.checkpoint import random_init_checkpoint, set_config
    from seq2squiggle_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version / debug lines must not land on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = set_config(None)
    sd = random_init_checkpoint(cfg, seed=1)["state_dict"]
    eng = Engine(sd, cfg, device=local)
    opts = default_opts(args.precision)
    lib = _lib.load()
    k = cfg["seq_kmer"]

    n_batches = args.warmup + args.steps
    # reads are sharded by rank: every rank simulates its own disjoint read set (weak scaling)
    host = [Engine.pack_reads(synth_reads(args.reads_per_step, seed=1000 * rank + i), k, pin=True) for i in range(n_batches)]
    devb = [(b.to(dev), ro.to(dev), co.to(dev), ro.numel() - 1, int(co[-1])) for b, ro, co in host]
    chunk_base = np.cumsum([0] + [x[4] for x in devb])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        b, ro, co, nr, nc = devb[i]
        return eng.forward_reads_device(b, ro, co, nr, nc, opts, chunk_id_base=int(chunk_base[i]) + rank * (1 << 40))

    for i in range(args.warmup):
        step(i)
    eng.check()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    l0 = lib.s2s_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    outs = []
    for i in range(args.warmup, n_batches):
        outs.append(step(i)[1])          # keep raw_offsets to count the emitted samples afterwards
    ev1.record()
    barrier()
    launches = lib.s2s_launch_count() - l0
    eng.check()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    samples = sum(int(o[-1]) for o in outs)
    chunks = sum(devb[i][4] for i in range(args.warmup, n_batches))
    reads = sum(devb[i][3] for i in range(args.warmup, n_batches))

    # ---- end to end through the public API with HOST inputs: model.predict_reads() = pack + pinned H2D + hot path +
    # D2H of offsets and int16 signal into host memory, copies overlapped with the next batch's compute
    from seq2squiggle_b200.model import seq2squiggle
    from seq2squiggle_b200.profiles import get_profile

    class _Sink:                                   # writer plug point that only counts (no file I/O in the metric)
        profile, profile_name = get_profile("dna-r10-prom"), "dna-r10-prom"
        signals, samples = None, 0

        def save(self):
            self.samples += sum(len(v) for v in self.signals.values())

    sink = _Sink()
    model = seq2squiggle(config=cfg, state_dict=sd, out_writer=sink, dwell_mean=12.5, dwell_std=0.0, noise_std=2.0,
                         noise_sampling=True, duration_sampling=True, min_noise=0.0, min_duration=3, device=local,
                         seed=7, precision=args.precision)
    host_reads = [[(r.decode("latin-1"), str(j)) for j, r in enumerate(synth_reads(args.reads_per_step, seed=1000 * rank + i))]
                  for i in range(n_batches)]
    for rd in host_reads[:args.warmup]:            # W untimed warm-up steps (workspace, pinned pools, allocator caches)
        model.predict_reads(rd)
    model.on_predict_epoch_end()
    sink.samples = 0
    stats0 = dict(model._pipe.stats)               # the pipeline (and its counters) lives as long as the model
    barrier()
    t0 = time.perf_counter()
    e2e_marks = [t0]
    for rd in host_reads[args.warmup:]:
        model.predict_reads(rd)
        e2e_marks.append(time.perf_counter())
    model.on_predict_epoch_end()
    pipe_stats = {k: v - stats0.get(k, 0) for k, v in model._pipe.stats.items()}
    torch.cuda.synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    if rank == 0:
        sys.stderr.write("e2e per-step submit ms: " + " ".join(f"{1e3 * (b - a):.0f}" for a, b in zip(e2e_marks, e2e_marks[1:]))
                         + f"; drain {1e3 * (t0 + e2e_s - e2e_marks[-1]):.0f} ms; staging allocations in the timed region: "
                         + f"{pipe_stats.get('allocs', 0)}\n")
    e2e_samples, h2d, d2h = sink.samples, pipe_stats["h2d_bytes"], pipe_stats["d2h_bytes"]
    eng.check()

    # ---- per-kernel timing of the dominant kernel (attention) for the roofline, outside the timed region
    kt = kernel_timing(eng, lib, step, args.warmup) if args.precision == "fp16" else None

    stats = torch.tensor([ms, e2e_s, samples, chunks, reads, e2e_samples, launches, h2d, d2h], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_s = float(mx[0]), float(mx[1])
        samples, chunks, reads, e2e_samples, launches, h2d, d2h = [float(x) for x in sm[2:]]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    value = samples / (ms * 1e-3)
    line = {"metric": "simulated raw-signal samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16" if args.precision == "fp16" else "f32",
            "data": "synthetic", "config": workload_config(args.reads_per_step, int(os.environ.get("S2S_BATCH_CHUNKS", "0")) or 32768),
            "reads_per_s": reads / (ms * 1e-3), "chunks_per_s": chunks / (ms * 1e-3),
            "decoder_positions_per_s": 250 * chunks / (ms * 1e-3),
            "model_tflops": FLOP_PER_CHUNK * chunks / (ms * 1e-3) / 1e12,
            "e2e": {"value": e2e_samples / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": h2d / args.steps / world,
                    "d2h_bytes_per_step": d2h / args.steps / world},
            "gpu_launches": int(launches), "clocks": clk}
    if kt:
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        ach = ATT_FLOP_PER_CHUNK_LAYER * kt["chunks_per_launch"] / (kt["ms_per_launch"] * 1e-3) / 1e12
        sm_mhz = (clk or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        exp_rate = ATT_EXP_PER_CHUNK_LAYER * kt["chunks_per_launch"] / (kt["ms_per_launch"] * 1e-3)
        exp_peak = MUFU_PER_CLK_SM * 148 * sm_mhz * 1e6
        line["roofline"] = {"kernel": "k_tc_attn2 (fused QKV projection + decoder attention)", "bound": "tensor",
                            "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                            "traffic": ATT_DRAM_BYTES_PER_CHUNK_NCU * kt["chunks_per_launch"],
                            "traffic_note": "dram__bytes_read+write per launch from the ncu --set full capture "
                                            "(profiles/r01_attn2_ncu.txt), scaled to this run's chunks per launch; "
                                            f"algorithmic HBM bytes per launch {ATT_ALGO_BYTES_PER_CHUNK * kt['chunks_per_launch']:.4g}",
                            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                            "launches_timed": kt["launches"], "ms_per_launch": kt["ms_per_launch"],
                            "share_of_step": kt["share"],
                            "note": "d_k=8 attention is bound by the softmax exponentials (XU pipe: MUFU.EX2 + F2FP), not by "
                                    "the tensor pipe; achieved counts attention MMA FLOPs only (QK^T + PV, 16 MFLOP per "
                                    "chunk per layer); exp.achieved counts all exponentials although a quarter of "
                                    "them run on the FMA pipe",
                            "exp": {"achieved_gexp_s": exp_rate / 1e9, "peak_gexp_s": exp_peak / 1e9,
                                    "frac": exp_rate / exp_peak, "peak": f"16 ex2/clk/SM x 148 SMs x {sm_mhz:.0f} MHz"}}
    if args.cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        rd = synth_reads(args.ref_reads, seed=99)
        cpu_reference_step(sd, cfg, rd[:8])
        n, c, t = cpu_reference_step(sd, cfg, rd)
        line["cpu_baseline"] = {"value": n / t, "unit": "samples/s", "cores": cores, "kind": "port",
                                "sample": f"{args.ref_reads} reads ({c} chunks) of the same distribution, oracle port "
                                          f"of the reference CPU path incl. its Python tokeniser, {t:.1f} s"}
    emit_result(line)
    if world > 1:
        dist.destroy_process_group()


def kernel_timing(eng, lib, step, i):
    """CUDA-event time of every k_tc_attention launch of one step (the library brackets the launches itself)."""
    import ctypes as C
    import torch
    if not hasattr(lib, "s2s_profile_kernel"):
        return None
    lib.s2s_profile_kernel.restype = C.c_int
    lib.s2s_profile_kernel.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    ms, n, ch = C.c_double(), C.c_int64(), C.c_int64()
    lib.s2s_profile_kernel(eng.handle, 1, None, None, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step(i)
    e1.record()
    torch.cuda.synchronize()
    lib.s2s_profile_kernel(eng.handle, 0, C.byref(ms), C.byref(n), C.byref(ch))
    if n.value == 0:
        return None
    return {"ms_per_launch": ms.value / n.value, "launches": int(n.value), "chunks_per_launch": ch.value / n.value,
            "share": ms.value / e0.elapsed_time(e1)}


_RESULT_FD = None


def emit_result(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _RESULT_FD is None:
        os.write(1, data)
    else:
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--precision", choices=["fp16", "fp32"], default="fp16")
    ap.add_argument("--reads-per-step", type=int, default=4000)
    ap.add_argument("--ref-reads", type=int, default=200, help="reads per CPU-arm step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # stdout carries exactly ONE line, the JSON result.  Libraries write there too (NCCL prints "NCCL version ..." to
    # stdout whatever NCCL_DEBUG_FILE says), so file descriptor 1 is pointed at stderr for the whole run and the result
    # line goes to the saved descriptor.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
