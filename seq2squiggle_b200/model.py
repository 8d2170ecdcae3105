"""The model plug point of ``seq2squiggle predict`` (reference ``model.py:25-63, 195-307``) on the B200 engine.

``seq2squiggle.load_from_checkpoint(ckpt, out_writer=..., dwell_mean=..., ...)`` returns an object with the
reference's predict surface — ``predict_step(batch)``, ``export_and_clear_results(keep_last)``,
``on_predict_epoch_end()``, ``.hparams.config`` — whose arithmetic runs entirely in ``libs2s_b200.so``:

* ``predict_step((read_ids, one_hot[B,16,k,5]))`` is the DataLoader-batch form (``s2s_forward_chunks``); rows are
  kept on the device and grouped per read at export time, where zero-strip + digitisation + per-read compaction
  are one CUDA pass (``s2s_compact_reads``) instead of a ``nonzero()`` sync and a NumPy round per read;
* ``predict_reads([(sequence, name), ...])`` is the native fast path used by ``inference_run``: read bytes go to
  the GPU, tokenisation happens there, int16 signals come back (``s2s_forward_reads``), with the device->host
  copies and the file writer overlapped with the next batch's compute on a side stream / writer thread.

No Lightning: the checkpoint is a plain ``torch.load`` of the Lightning file layout (``checkpoint.py``).
"""
from __future__ import annotations

import logging
import os
import queue
import threading
import time
from collections import OrderedDict
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from .checkpoint import load_checkpoint
from .engine import Engine, RunOptions
from .signal_io import BLOW5Writer

logger = logging.getLogger("seq2squiggle")

PIPE_TRACE = bool(int(os.environ.get("S2S_PIPE_TRACE", "0")))
PIPE_CHUNKS = int(os.environ.get("S2S_PIPE_CHUNKS", 65536))   # chunks per pipeline piece of predict_reads (2 engine sub-batches)
PIPE_DEPTH = 3        # pieces the host may queue ahead of the one whose result it waits for (absorbs host jitter)
PIPE_SLOTS = PIPE_DEPTH + 6   # pinned staging slots: PIPE_DEPTH + 1 on the device side, 4 queued for the writer, 1 in save()


class _HParams(dict):
    """``model.hparams`` as Lightning exposes it: attribute and item access."""
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


class seq2squiggle:
    """Feed-forward-transformer signal predictor; same keyword-only init arguments as model.py:30-44."""

    def __init__(self, *, config: dict, save_valid_plots: bool = True, out_writer=None, dwell_mean: float = 9.0,
                 dwell_std: float = 0.0, noise_std: float = -1, noise_sampling: bool = False,
                 duration_sampling: bool = False, export_every_n_samples: int = 2000000, min_noise: float = 0.5,
                 min_duration: int = 1, state_dict: Optional[Dict[str, torch.Tensor]] = None, device: int = 0,
                 seed: Optional[int] = None, precision: str = "fp16", profile: Optional[dict] = None,
                 profile_name: Optional[str] = None):
        if state_dict is None:
            raise ValueError("seq2squiggle_b200 is inference-only: construct it with load_from_checkpoint() or pass "
                             "state_dict=")
        self.hparams = _HParams(config=config, save_valid_plots=save_valid_plots, out_writer=out_writer,
                                dwell_mean=dwell_mean, dwell_std=dwell_std, noise_std=noise_std,
                                noise_sampling=noise_sampling, duration_sampling=duration_sampling,
                                export_every_n_samples=export_every_n_samples, min_noise=min_noise,
                                min_duration=min_duration)
        self.config = config
        self.save_valid_plots = save_valid_plots
        self.results: list = []
        self.out_writer = out_writer
        self.dwell_mean, self.dwell_std = dwell_mean, dwell_std
        self.noise_std, self.noise_sampling = noise_std, noise_sampling
        self.duration_sampling = duration_sampling
        self.export_every_n_samples = export_every_n_samples
        self.total_samples = 0
        self.min_noise, self.min_duration = min_noise, min_duration
        self.precision = precision
        self.engine = Engine(state_dict, config, device=device)
        self.device = self.engine.device
        # The reference draws from torch's global generator seeded by set_seeds (utils.py:722-741); the device
        # Philox streams are keyed by that same seed.
        self.seed = int(torch.initial_seed() if seed is None else seed)
        self.chunks_done = 0          # global chunk index base of the next batch (Philox counter)
        if profile is None and out_writer is not None:
            profile, profile_name = out_writer.profile, out_writer.profile_name
        self._profile, self._profile_name = profile, profile_name
        self._pipe: Optional[_ReadPipeline] = None

    # ------------------------------------------------------------------------------------------
    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
        """inference.py:386-397: keyword arguments override the checkpoint's hyper-parameters."""
        sd, hp = load_checkpoint(str(checkpoint_path))
        init = {k: v for k, v in hp.items() if k in ("config", "save_valid_plots", "out_writer", "dwell_mean", "dwell_std",
                                                    "noise_std", "noise_sampling", "duration_sampling",
                                                    "export_every_n_samples", "min_noise", "min_duration")}
        init.update(kwargs)
        return cls(state_dict=sd, **init)

    def eval(self):
        return self

    def run_options(self) -> RunOptions:
        p = self._profile or {}
        return RunOptions(dwell_mean=float(self.dwell_mean), dwell_std=float(self.dwell_std),
                          duration_sampling=bool(self.duration_sampling), min_duration=float(self.min_duration),
                          noise_std=float(self.noise_std), noise_sampling=bool(self.noise_sampling),
                          min_noise=float(self.min_noise), digitisation=float(p.get("digitisation", 2048.0)),
                          range=float(p.get("range", 281.345551)), offset_mean=float(p.get("offset_mean", -127.5655735)),
                          rna=bool(self._profile_name and self._profile_name.startswith("rna")), seed=self.seed,
                          precision=self.precision)

    # ------------------------------------------------------------------------------------------
    # DataLoader-batch plug point (model.py:195-250)
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def predict_step(self, batch):
        read_id, data, *_ = batch
        bs, seq_l = data.shape[:2]
        k = self.config["seq_kmer"]
        oh = data.to(self.device, non_blocking=True).reshape(bs, seq_l, k, 5)
        # one-hot -> letter code (argmax; an all-zero row, i.e. a letter outside "_ACGT", is -1)
        codes = torch.where(oh.amax(-1) > 0, oh.argmax(-1), torch.full((), -1, device=self.device)).to(torch.int8)
        pa, _ = self.engine.forward_chunks(codes.contiguous(), self.run_options(), chunk_id_base=self.chunks_done,
                                           check=False)
        self.chunks_done += bs
        self.results.append((list(read_id), pa))
        self.total_samples += bs                                   # model.py:247 counts chunks
        if isinstance(self.out_writer, BLOW5Writer) and self.total_samples >= self.export_every_n_samples:
            self.export_and_clear_results(keep_last=True)
            self.total_samples = 0

    def export_and_clear_results(self, keep_last: bool = True):
        """model.py:253-302: group rows per read (first-seen order), hold back the last read if ``keep_last``, strip
        exact zeros, hand ``{read_id: signal}`` to the writer."""
        order: "OrderedDict[str, list]" = OrderedDict()
        row = 0
        for ids, _ in self.results:
            for rid in ids:
                order.setdefault(rid, []).append(row)
                row += 1
        pa_all = torch.cat([pa for _, pa in self.results]) if self.results else None
        last = None
        if keep_last and order:
            last_key = next(reversed(order))
            last = (last_key, order.pop(last_key))
        signals: "OrderedDict[str, np.ndarray]" = OrderedDict()
        if order:
            rows = torch.tensor([r for v in order.values() for r in v], dtype=torch.int64, device=self.device)
            counts = np.fromiter((len(v) for v in order.values()), dtype=np.int64, count=len(order))
            chunk_off = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)).to(self.device)
            raw, raw_off = self.engine.compact_reads(pa_all.index_select(0, rows), chunk_off, self.run_options())
            self.engine.check()
            off = raw_off.cpu().numpy()
            sig = raw[: int(off[-1])].cpu().numpy()
            for i, rid in enumerate(order):
                signals[rid] = sig[off[i]:off[i + 1]]
        self.out_writer.signals = signals
        self.out_writer.save()
        self.out_writer.signals = []
        self.results = []
        if last is not None:
            self.results.append(([last[0]] * len(last[1]),
                                 pa_all.index_select(0, torch.tensor(last[1], dtype=torch.int64, device=self.device))))
        logger.debug("Results exported and memory cleared.")

    def on_predict_epoch_end(self):
        if self._pipe is not None:
            self._pipe.finish()          # drained; streams and staging slots stay for the next epoch
        if self.results:
            self.export_and_clear_results(keep_last=False)
        logger.debug("Epoch end operation completed.")

    # ------------------------------------------------------------------------------------------
    # native fast path: whole reads in, int16 signals out, copies and writer overlapped with compute
    # ------------------------------------------------------------------------------------------
    def predict_reads(self, reads: Sequence[Tuple[str, str]], chunk_id_base: Optional[int] = None, tag=None):
        """``reads``: ``[(sequence, name), ...]`` (what ``get_reads`` yields).  Every read is complete, so each batch
        is exported as soon as its signal reaches the host (no ``keep_last`` hold-back needed).  ``tag`` (sharded runs:
        ``(batch index, first global read)``) makes the whole call ONE pipeline piece and is handed to the writer."""
        if self._pipe is None:
            self._pipe = _ReadPipeline(self)
        # pieces of about PIPE_CHUNKS chunks: the copy / writer stages run one piece behind the compute stage, so the
        # un-overlapped tail of a run is one piece, not one caller-sized batch
        k, piece, n = self.engine.k, [], 0
        base = chunk_id_base
        for item in reads:
            piece.append(item)
            nk = len(item[0]) - k + 1
            n += (nk + 15) // 16 if nk > 0 else 0
            if n >= PIPE_CHUNKS and tag is None:
                self._pipe.submit(piece, base)
                base = None if base is None else base + n
                piece, n = [], 0
        if piece or tag is not None:
            self._pipe.submit(piece, base, tag)


class _Slot:
    """Staging of one pipeline piece: pinned host buffers (bases + offsets in, offsets + int16 signal out) and the
    device buffers they are copied to / from.  Slots are recycled: in steady state NOTHING is allocated, neither pinned
    memory (cudaHostAlloc / cudaFreeHost synchronise the whole device) nor device memory (a cudaMalloc on a busy device
    stalled the enqueueing thread for up to 0.5 s in traces) — either cost the host its lead over the GPU at random."""

    def __init__(self, device):
        self.device = device
        self.bases = self.ro = self.co = self.off = self.sig = None          # pinned host
        self.bases_d = self.ro_d = self.co_d = self.raw_d = self.off_d = None  # device

    @staticmethod
    def _grown(t, n, dtype, quantum, **kw):
        if t is not None and t.numel() >= n:
            return t, 0
        cap = -(-int(1.25 * n + 1) // quantum) * quantum
        return torch.empty(cap, dtype=dtype, **kw), 1

    def fit(self, n_bases, n_reads, n_chunks) -> int:
        """Make every buffer large enough for such a piece; returns the number of (re)allocations."""
        pin, dev = dict(pin_memory=True), dict(device=self.device)
        n_b, n_r, n_s = max(n_bases, 1), n_reads + 1, max(n_chunks * 250, 1)   # 250 samples per chunk at most
        grew = 0
        for name, n, dt, q, kw in (("bases", n_b, torch.uint8, 1 << 20, pin), ("ro", n_r, torch.int64, 1 << 12, pin),
                                   ("co", n_r, torch.int64, 1 << 12, pin), ("off", n_r, torch.int64, 1 << 12, pin),
                                   ("sig", n_s, torch.int16, 8 << 20, pin), ("bases_d", n_b, torch.uint8, 1 << 20, dev),
                                   ("ro_d", n_r, torch.int64, 1 << 12, dev), ("co_d", n_r, torch.int64, 1 << 12, dev),
                                   ("off_d", n_r, torch.int64, 1 << 12, dev), ("raw_d", n_s, torch.int16, 8 << 20, dev)):
            t, g = self._grown(getattr(self, name), n, dt, q, **kw)
            setattr(self, name, t)
            grew += g
        return grew


class _ReadPipeline:
    """compute stream: H2D(bases) -> s2s_forward_reads;  copy stream: D2H(offsets) -> D2H(int16 prefix);  writer
    thread: writer.signals = {...}; writer.save().  PIPE_DEPTH pieces of compute are queued ahead of the copies.
    One pipeline (streams, staging slots) lives as long as its model; ``finish()`` drains it at the end of an epoch."""

    def __init__(self, model: seq2squiggle):
        self.m = model
        self.eng = model.engine
        self.dev = model.device
        self.compute = torch.cuda.Stream(self.dev)
        self.copy = torch.cuda.Stream(self.dev)
        self.inflight: list = []     # pieces whose compute is queued but whose signal has not been fetched
        self.q: "queue.Queue" = queue.Queue(maxsize=4)
        self.err: Optional[BaseException] = None
        self.stats = dict(reads=0, chunks=0, samples=0, h2d_bytes=0, d2h_bytes=0, allocs=0)
        self.trace: list = []        # S2S_PIPE_TRACE=1: per-piece host timestamps and device events (developer aid)
        self.free_slots: "queue.Queue" = queue.Queue()
        self.pending: "OrderedDict" = OrderedDict()   # reads kept for a writer that can only write once (POD5)
        self.n_slots = 0
        self.thread: Optional[threading.Thread] = None

    def _ensure_writer(self):
        if self.thread is None or not self.thread.is_alive():
            self.thread = threading.Thread(target=self._writer_loop, daemon=True)
            self.thread.start()

    def _slot(self, n_bases, n_reads, n_chunks) -> _Slot:
        """A free staging slot fitted to the piece.  The whole pool of PIPE_SLOTS slots (in flight on the device +
        queued for the writer + being written) is created on the first piece, sized for pieces like it; afterwards the
        host waits for the writer thread to hand a slot back instead of allocating."""
        if self.n_slots == 0:
            with torch.cuda.stream(self.compute):
                for _ in range(PIPE_SLOTS):
                    sl = _Slot(self.dev)
                    sl.fit(n_bases, n_reads, n_chunks)
                    self.free_slots.put(sl)
            self.n_slots = PIPE_SLOTS
        while True:
            if self.err:
                raise self.err
            try:
                sl = self.free_slots.get(timeout=0.5)
                break
            except queue.Empty:
                continue
        with torch.cuda.stream(self.compute):
            self.stats["allocs"] += sl.fit(n_bases, n_reads, n_chunks)
        return sl

    def submit(self, reads, chunk_id_base=None, tag=None):
        if self.err:
            raise self.err
        self._ensure_writer()
        m = self.m
        t_sub = time.perf_counter()
        names = [n for _, n in reads]
        joined, read_off, chunk_off = Engine.pack_reads_np([s for s, _ in reads], self.eng.k)
        n_reads, n_chunks, n_bases = len(names), int(chunk_off[-1]), len(joined)
        slot = self._slot(n_bases, n_reads, n_chunks)
        slot.bases[:n_bases].numpy()[:] = np.frombuffer(joined, dtype=np.uint8)
        slot.ro[:n_reads + 1].numpy()[:] = read_off
        slot.co[:n_reads + 1].numpy()[:] = chunk_off
        base = m.chunks_done if chunk_id_base is None else chunk_id_base
        m.chunks_done = base + n_chunks
        t_packed = time.perf_counter()
        nb = max(n_bases, 1)
        with torch.cuda.stream(self.compute):
            ev_start = None
            if PIPE_TRACE:
                ev_start = torch.cuda.Event(enable_timing=True)
                ev_start.record(self.compute)
            slot.bases_d[:nb].copy_(slot.bases[:nb], non_blocking=True)
            slot.ro_d[:n_reads + 1].copy_(slot.ro[:n_reads + 1], non_blocking=True)
            slot.co_d[:n_reads + 1].copy_(slot.co[:n_reads + 1], non_blocking=True)
            raw, raw_off, _ = self.eng.forward_reads_device(slot.bases_d[:nb], slot.ro_d[:n_reads + 1],
                                                            slot.co_d[:n_reads + 1], n_reads, n_chunks, m.run_options(),
                                                            base, out=(slot.raw_d, slot.off_d[:n_reads + 1]))
            done = torch.cuda.Event(enable_timing=PIPE_TRACE)
            done.record(self.compute)
        off_host = slot.off[:n_reads + 1]
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(done)
            off_host.copy_(raw_off, non_blocking=True)
            off_ev = torch.cuda.Event()
            off_ev.record(self.copy)
        cur = dict(names=names, off_host=off_host, off_ev=off_ev, slot=slot, n_chunks=n_chunks, tag=tag)
        if PIPE_TRACE:
            cur["trace"] = dict(t_sub=t_sub, t_packed=t_packed, t_enq=time.perf_counter(), ev_start=ev_start, ev_done=done)
        self.stats["h2d_bytes"] += n_bases + 16 * (n_reads + 1)
        self.inflight.append(cur)
        while len(self.inflight) > PIPE_DEPTH:      # the host runs PIPE_DEPTH pieces ahead of the device
            self._fetch(self.inflight.pop(0))

    def _fetch(self, b):
        t_f0 = time.perf_counter()
        b["off_ev"].synchronize()
        if PIPE_TRACE:
            tr = b["trace"]
            tr.update(t_fetch0=t_f0, t_fetch1=time.perf_counter())
            self.trace.append(tr)
        n = int(b["off_host"][-1])
        slot = b["slot"]
        with torch.cuda.stream(self.copy):
            slot.sig[:n].copy_(slot.raw_d[:n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy)
        b.update(n=n, sig_ev=ev)
        self.stats["d2h_bytes"] += 2 * n + 8 * b["off_host"].numel()
        self.stats["reads"] += len(b["names"]); self.stats["chunks"] += b["n_chunks"]; self.stats["samples"] += n
        self.q.put(b)

    def _writer_loop(self):
        while True:
            b = self.q.get()
            if b is None:
                return
            try:
                if self.err is not None:    # an earlier piece failed: writing on would leave a gap in the output
                    continue
                b["sig_ev"].synchronize()
                off = b["off_host"].numpy()
                sig = b["slot"].sig.numpy()
                w = self.m.out_writer
                if w is not None and hasattr(w, "save_flat"):
                    if b.get("tag") is not None:           # sharded run: ordered write into the shared output file
                        w.save_flat(b["names"], sig, off, tag=b["tag"])
                    else:
                        w.save_flat(b["names"], sig, off)  # one contiguous buffer + offsets: no per-read Python work
                elif w is not None and not getattr(w, "appendable", True):
                    # a writer that cannot append (POD5: inference.py:71-79 sets export_every_n_samples = inf, all reads
                    # are kept and written once at on_predict_epoch_end): keep copies, the staging slot is recycled
                    for i, name in enumerate(b["names"]):
                        self.pending[name] = sig[off[i]:off[i + 1]].copy()
                elif w is not None:    # the views are valid until save() returns (writer plug-point contract)
                    w.signals = OrderedDict((name, sig[off[i]:off[i + 1]]) for i, name in enumerate(b["names"]))
                    w.save()
                    w.signals = []
            except BaseException as exc:  # surfaced on the next submit()/finish()
                self.err = exc
            finally:
                self.free_slots.put(b.pop("slot"))

    def finish(self):
        """Drain: fetch everything in flight, let the writer finish, surface errors.  The pipeline stays usable."""
        while self.inflight:
            self._fetch(self.inflight.pop(0))
        if self.thread is not None and self.thread.is_alive():
            self.q.put(None)
            self.thread.join()
        if self.pending and self.err is None:
            w = self.m.out_writer
            w.signals, self.pending = self.pending, OrderedDict()
            w.save()
            w.signals = []
        self.eng.check()
        if self.err:
            err, self.err = self.err, None
            raise err
