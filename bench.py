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
ATT_MUFU_SHARE = 10 / 16                    # k_tc_attn3: kPoly3H2 = 6 of 16 pairs by polynomial, 10 by MUFU.EX2
# k_tc_attn3, one launch of 32768 chunks, `ncu --set full` (profiles/r02_attn3_ncu.txt): dram__bytes_read.sum 1.074546 GB
# + dram__bytes_write.sum 1.040482 GB; the algorithmic traffic is 2 x 256 rows x 128 B = 65,536 B per chunk (x16 in, o16 out)
# (ncu's figure is just below it: the tail of the o16 stream is still in the 126 MB L2 when the counters stop)
ATT_DRAM_BYTES_PER_CHUNK_NCU = (2.154755e9 + 2.155325e9) / 65536   # profiles/r02_attn3_ncu.txt (one 65536-chunk launch)
ATT_ALGO_BYTES_PER_CHUNK = 2 * 256 * 128
ATT_QKV_FLOP_PER_CHUNK_LAYER = 2 * 256 * 64 * 192          # fused QKV projection inside the attention kernel
FFN_FLOP_PER_CHUNK_LAYER = 2 * 256 * 64 * (64 + 256 + 256)  # fc + W1 + W2 per decoder layer (256 rows per chunk)
FFN_BYTES_PER_CHUNK_LAYER = 256 * 384                       # o16 read + x16 read + x16 written, 128 B per row each
LR_BYTES_PER_CHUNK = 2048 + 64 + 64 + 32000 + 1000 + 4      # enc_out, sigma, dur in; features, sigma_ext, total out
LR_WRITE_SHARE = (32000 + 1000 + 4) / LR_BYTES_PER_CHUNK      # 0.94: the kernel is store-dominated
COMPACT_BYTES_PER_CHUNK_FIXED = 1000 + 4 + 8                # pA read, count read, offset written (+ 2 B per emitted sample)


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
CPU_MATMUL_PRECISION = "medium"   # what the reference sets (model.py:22); "highest" is the golden-vector mode


def cpu_reference_step(sd, cfg, reads, batch_chunks=1024):
    """One bounded sample of the reference CPU path: tokenise (utils.py:350-356, the reference's Python loops),
    predict_step per DataLoader batch of 1024 chunks (model.py:195-250), export + digitise.  Returns
    (emitted samples, chunks, seconds)."""
    import torch
    from oracle import s2s_oracle as orc
    from oracle.profiles_kat import PROFILES
    prof = PROFILES["dna-r10-prom"]
    torch.set_float32_matmul_precision(CPU_MATMUL_PRECISION)
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
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample,
                             "precision": f"fp32 tensors, torch.set_float32_matmul_precision('{CPU_MATMUL_PRECISION}') "
                                          "as the reference sets it (model.py:22)"},
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
    if args.sharp != 1.0:
        # trained-like regime: W_q / W_k of the decoder scaled (scores grow with the square), so that the fast attention
        # kernel's fp16 probabilities overflow and units fall back to the exact kernel
        for layer in range(cfg["decoder_layers"]):
            for name in ("w_qs", "w_ks"):
                for part in ("weight", "bias"):
                    key = f"decoders.layer_stack_FFT.{layer}.slf_attn.{name}.{part}"
                    sd[key] = sd[key] * args.sharp
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

    # ---- the same end-to-end loop with the real writer: every rank appends its records to a BLOW5 file (native writer,
    # uncompressed records) while the GPU works on the next pieces
    e2e_w = None
    if args.writer_e2e:
        import tempfile
        from seq2squiggle_b200.signal_io import BLOW5Writer
        wdir = tempfile.mkdtemp(prefix="s2s_bench_")
        wpath = os.path.join(wdir, f"rank{rank}.blow5")
        bw = BLOW5Writer(wpath, get_profile("dna-r10-prom"), False, "dna-r10-prom", False, record_compression="none")
        model.out_writer = bw
        model.predict_reads(host_reads[0])          # untimed: creates the file, warms the writer thread
        model.on_predict_epoch_end()
        w0 = bw.samples_written
        barrier()
        t0w = time.perf_counter()
        for rd in host_reads[args.warmup:]:
            model.predict_reads(rd)
        model.on_predict_epoch_end()
        torch.cuda.synchronize()
        barrier()
        e2e_w = (bw.samples_written - w0, time.perf_counter() - t0w, os.path.getsize(wpath))
        model.out_writer = sink
        try:
            os.remove(wpath)
            os.rmdir(wdir)
        except OSError:
            pass

    # ---- per-kernel timing of the dominant kernel (attention) for the roofline, outside the timed region
    kt = kernel_timing(eng, lib, step, args.warmup) if args.precision == "fp16" else None

    ew = e2e_w or (0, 0.0, 0)
    stats = torch.tensor([ms, e2e_s, samples, chunks, reads, e2e_samples, launches, h2d, d2h, ew[0], ew[1], ew[2]],
                         dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_s = float(mx[0]), float(mx[1])
        samples, chunks, reads, e2e_samples, launches, h2d, d2h = [float(x) for x in sm[2:9]]
        ew = (float(sm[9]), float(mx[10]), float(sm[11]))
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
            "data": "synthetic", "config": workload_config(args.reads_per_step, int(os.environ.get("S2S_BATCH_CHUNKS", "0")) or 65536),
            "reads_per_s": reads / (ms * 1e-3), "chunks_per_s": chunks / (ms * 1e-3),
            "decoder_positions_per_s": 250 * chunks / (ms * 1e-3),
            "model_tflops": FLOP_PER_CHUNK * chunks / (ms * 1e-3) / 1e12,
            "e2e": {"value": e2e_samples / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": h2d / args.steps / world,
                    "d2h_bytes_per_step": d2h / args.steps / world},
            "gpu_launches": int(launches), "clocks": clk}
    if e2e_w:
        line["e2e_with_writer"] = {"value": ew[0] / ew[1], "unit": "samples/s", "file_bytes": ew[2],
                                   "sink": "one uncompressed BLOW5 file per rank in the temp directory, native writer "
                                           "thread overlapped with compute (the sharded CLI writes one shared file "
                                           "the same way: profiles/r02_config5_*gpu.txt)"}
    if args.sharp != 1.0:
        line["config"]["sharp_scale"] = args.sharp
    if kt:
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
        sm_mhz = (clk or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        att = kt["attention"]
        # write-only HBM rate of this GPU, measured here (the driver's peak is a COPY rate; a store-dominated kernel such as
        # the length regulator -- 94 % of its bytes are writes -- is bounded by this one): fill of a 2 GiB buffer, best of 5
        wbuf = torch.empty(1 << 31, dtype=torch.uint8, device=dev)
        w_ms = []
        for _ in range(6):
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record(); wbuf.fill_(1); w1.record(); torch.cuda.synchronize()
            w_ms.append(w0.elapsed_time(w1))
        hbm_write_peak = wbuf.numel() / (min(w_ms[1:]) * 1e-3) / 1e9
        del wbuf
        cpl, sec = att["chunks_per_launch"], att["ms_per_launch"] * 1e-3
        exp_rate = ATT_EXP_PER_CHUNK_LAYER * cpl / sec
        exp_peak = MUFU_PER_CLK_SM * 148 * sm_mhz * 1e6
        tens = (ATT_FLOP_PER_CHUNK_LAYER + ATT_QKV_FLOP_PER_CHUNK_LAYER) * cpl / sec / 1e12
        # The dominant kernel: d_k = 8 gives 32 MMA FLOPs per softmax element, so the kernel is bound by the exponentials
        # (XU pipe: MUFU.EX2 at 16 / clk / SM, shared with the F2FP packs), not by the tensor pipe.  `achieved` / `peak`
        # are the exponential rate against the MUFU-only rate (part of the exponentials run as a packed-fp16 polynomial
        # on the FMA pipe, so the fraction can pass 1); the tensor-pipe fraction of the same launches is alongside.
        line["roofline"] = {"kernel": "k_tc_attn3 (fused QKV projection + decoder attention, one launch = one decoder layer "
                                      "of one sub-batch of up to 65536 chunks, incl. k_attn_gate and the exact-kernel fallback pass; "
                                      "ms_per_32768_chunks is the figure earlier rounds quoted)",
                            "bound": "xu", "achieved": exp_rate / 1e9, "peak": exp_peak / 1e9, "unit": "Gexp/s",
                            "frac": exp_rate / exp_peak,
                            # 10 of every 16 exponential pairs go through MUFU.EX2 (the other 6 are the packed-fp16 polynomial on
                            # the FMA pipe): this is how busy the MUFU unit itself is; ncu's XU-pipe figure (73 %,
                            # profiles/r02_attn3_ncu.txt) also counts the F2FP conversions that share the pipe
                            "mufu_share_of_exp": ATT_MUFU_SHARE, "mufu_busy_frac": ATT_MUFU_SHARE * exp_rate / exp_peak,
                            "limiter": "issue slots (75 %) and XU pipe (73 %) together, ncu; four softmax warps per scheduler",
                            "peak_source": f"16 MUFU.EX2 / clk / SM x 148 SMs x {sm_mhz:.0f} MHz (SM clock sampled in this run)",
                            "tensor": {"achieved": tens, "peak": tf_peak, "unit": "TFLOP/s", "frac": tens / tf_peak,
                                       "peak_source": f"{src} bf16_tflops_sustained",
                                       "note": "QK^T + PV (16.0 MFLOP) + fused QKV projection (6.29 MFLOP) per chunk per layer"},
                            "traffic": ATT_DRAM_BYTES_PER_CHUNK_NCU * cpl,
                            "traffic_note": "dram__bytes_read+write per launch from the ncu --set full capture "
                                            "(profiles/r02_attn3_ncu.txt), scaled to this run's chunks per launch; "
                                            f"algorithmic HBM bytes per launch {ATT_ALGO_BYTES_PER_CHUNK * cpl:.4g}",
                            "launches_timed": att["launches"], "ms_per_launch": att["ms_per_launch"],
                            "chunks_per_launch": cpl, "ms_per_32768_chunks": att["ms_per_launch"] * 32768 / cpl,
                            "share_of_step": att["share"]}
        kernels = []
        if "ffn" in kt:
            f = kt["ffn"]
            ach = FFN_FLOP_PER_CHUNK_LAYER * f["chunks_per_launch"] / (f["ms_per_launch"] * 1e-3) / 1e12
            gbs = FFN_BYTES_PER_CHUNK_LAYER * f["chunks_per_launch"] / (f["ms_per_launch"] * 1e-3) / 1e9
            kernels.append({"kernel": "k_tc_fc_ffn4 (fc + LN + FFN + LN; last layer: + out_linear, x165, noise, clamp, count)",
                            "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                            "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak,
                            "ms_per_launch": f["ms_per_launch"], "launches_timed": f["launches"], "share_of_step": f["share"]})
            kernels[-1]["ms_per_32768_chunks"] = f["ms_per_launch"] * 32768 / f["chunks_per_launch"]
        if "length_regulate" in kt:
            f = kt["length_regulate"]
            ach = LR_BYTES_PER_CHUNK * f["chunks_per_launch"] / (f["ms_per_launch"] * 1e-3) / 1e9
            kernels.append({"kernel": "k_length_regulate16", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                            "frac": ach / hbm_peak, "write_share_of_bytes": LR_WRITE_SHARE,
                            "frac_of_write_only_peak": ach / hbm_write_peak, "ms_per_launch": f["ms_per_launch"],
                            "launches_timed": f["launches"], "share_of_step": f["share"]})
            kernels[-1]["ms_per_32768_chunks"] = f["ms_per_launch"] * 32768 / f["chunks_per_launch"]
        if "compact" in kt:
            f = kt["compact"]
            byt = (COMPACT_BYTES_PER_CHUNK_FIXED + 2.0 * samples / max(chunks, 1)) * f["chunks_per_launch"]
            ach = byt / (f["ms_per_launch"] * 1e-3) / 1e9
            kernels.append({"kernel": "zero-strip compaction + digitisation (scan x3, k_read_offsets, k_compact)",
                            "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                            "ms_per_launch": f["ms_per_launch"], "launches_timed": f["launches"], "share_of_step": f["share"]})
        for name in ("encoder", "front_end"):
            if name in kt:
                f = kt[name]
                kernels.append({"kernel": name, "ms_per_launch": f["ms_per_launch"], "launches_timed": f["launches"],
                                "ms_per_32768_chunks": f["ms_per_launch"] * 32768 / f["chunks_per_launch"],
                                "share_of_step": f["share"]})
        line["kernels"] = kernels
        line["peaks"] = {"hbm_gbs": hbm_peak, "bf16_tflops_sustained": tf_peak, "source": src,
                         "hbm_write_only_gbs": hbm_write_peak,
                         "hbm_write_only_source": "torch fill_ of 2 GiB timed in this run (best of 5)"}
    if args.cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        rd = synth_reads(args.ref_reads, seed=99)
        cpu_reference_step(sd, cfg, rd[:8])
        n, c, t = cpu_reference_step(sd, cfg, rd)
        line["cpu_baseline"] = {"value": n / t, "unit": "samples/s", "cores": cores, "kind": "port",
                                "sample": f"{args.ref_reads} reads ({c} chunks) of the same distribution, oracle port "
                                          f"of the reference CPU path incl. its Python tokeniser, {t:.1f} s",
                                "precision": f"fp32 tensors, torch.set_float32_matmul_precision('{CPU_MATMUL_PRECISION}') "
                                             "as the reference sets it (model.py:22)"}
    if args.gpu_eager_baseline and world == 1:
        line["gpu_eager_baseline"] = gpu_eager_baseline(sd, cfg, dev, samples / max(chunks, 1))
    emit_result(line)
    if world > 1:
        dist.destroy_process_group()


def kernel_timing(eng, lib, step, i):
    """CUDA-event time of the kernel groups of one step (the library brackets its own launches on the launching
    stream, s2s_profile_kernel / s2s_profile_kernel_group)."""
    import ctypes as C
    import torch
    if not hasattr(lib, "s2s_profile_kernel_group"):
        return None
    sig = [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.s2s_profile_kernel.restype = C.c_int
    lib.s2s_profile_kernel.argtypes = [C.c_void_p, C.c_int] + sig
    lib.s2s_profile_kernel_group.restype = C.c_int
    lib.s2s_profile_kernel_group.argtypes = [C.c_void_p, C.c_char_p] + sig
    lib.s2s_profile_kernel(eng.handle, 1, None, None, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step(i)
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1)
    out = {}
    for name in ("attention", "ffn", "length_regulate", "compact", "encoder", "front_end"):
        ms, n, ch = C.c_double(), C.c_int64(), C.c_int64()
        if lib.s2s_profile_kernel_group(eng.handle, name.encode(), C.byref(ms), C.byref(n), C.byref(ch)) != 0 or n.value == 0:
            continue
        out[name] = {"ms_per_launch": ms.value / n.value, "launches": int(n.value), "chunks_per_launch": ch.value / n.value,
                     "share": ms.value / step_ms}
    lib.s2s_profile_kernel(eng.handle, 0, None, None, None)
    return out if "attention" in out else None


def gpu_eager_baseline(sd, cfg, dev, samples_per_chunk):
    """The reference's own GPU mode (inference.py:404: Lightning precision "16-mixed"): the oracle's encoder + length
    regulator + decoder in eager PyTorch under fp16 autocast on this GPU, DataLoader batches of 1024 chunks
    (--predict-batch-size default), model only (no tokeniser, no samplers, no writer): the number the kernels must beat."""
    import torch
    from oracle import s2s_oracle as orc
    k = cfg["seq_kmer"]
    g = torch.Generator().manual_seed(3)
    codes = torch.randint(1, 5, (1024, 16, k), generator=g)
    data = torch.nn.functional.one_hot(codes, 5).to(torch.float16).reshape(1024, 16, 5 * k).to(dev)
    dur = np.full((1024, 16), 12, dtype=np.int32)
    j = torch.from_numpy(orc.lr_expand_indices(dur, 250)).to(dev).long()
    sd_dev = {key: v.to(dev) for key, v in sd.items() if torch.is_tensor(v)}

    def eager(x):
        with torch.inference_mode(), torch.autocast("cuda", dtype=torch.float16):
            enc, _ = orc.encoder_forward(sd_dev, cfg, x)
            lr = torch.where(j[..., None] >= 0, torch.gather(enc.float(), 1, j.clamp(min=0)[..., None].expand(-1, -1, 64)), 0.0)
            return orc.decoder_forward(sd_dev, cfg, lr).squeeze(-1).float() * 165.0

    for _ in range(3):
        eager(data)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_it = 20
    e0.record()
    for _ in range(n_it):
        eager(data)
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3 / n_it
    return {"what": "oracle encoder + length regulator + decoder, eager PyTorch under torch.autocast(float16) on this GPU "
                    "(the reference's GPU mode, inference.py:404), 1024-chunk batches, model only",
            "ms_per_batch": dt * 1e3, "chunks_per_s": 1024 / dt, "value": 1024 / dt * samples_per_chunk, "unit": "samples/s",
            "note": "samples/s = chunks/s x this run's emitted samples per chunk"}


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
    ap.add_argument("--no-gpu-eager-baseline", dest="gpu_eager_baseline", action="store_false",
                    help="skip timing the reference's own GPU mode (oracle modules, eager PyTorch, fp16 autocast)")
    ap.add_argument("--no-writer-e2e", dest="writer_e2e", action="store_false",
                    help="skip the second end-to-end loop that writes BLOW5 files")
    ap.add_argument("--sharp", type=float, default=1.0,
                    help="scale W_q / W_k of the decoder (trained-like sharp attention: exercises the exact-kernel fallback)")
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
