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
import queue
import threading
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .checkpoint import load_checkpoint
from .engine import Engine, RunOptions
from .signal_io import BLOW5Writer

logger = logging.getLogger("seq2squiggle")

PIPE_CHUNKS = 65536   # chunks per pipeline piece of predict_reads (2 engine sub-batches)
PIPE_DEPTH = 3        # pieces the host may queue ahead of the one whose result it waits for (absorbs host jitter)


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
            self._pipe.finish()
            self._pipe = None
        if self.results:
            self.export_and_clear_results(keep_last=False)
        logger.debug("Epoch end operation completed.")

    # ------------------------------------------------------------------------------------------
    # native fast path: whole reads in, int16 signals out, copies and writer overlapped with compute
    # ------------------------------------------------------------------------------------------
    def predict_reads(self, reads: Sequence[Tuple[str, str]], chunk_id_base: Optional[int] = None):
        """``reads``: ``[(sequence, name), ...]`` (what ``get_reads`` yields).  Every read is complete, so each batch
        is exported as soon as its signal reaches the host (no ``keep_last`` hold-back needed)."""
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
            if n >= PIPE_CHUNKS:
                self._pipe.submit(piece, base)
                base = None if base is None else base + n
                piece, n = [], 0
        if piece:
            self._pipe.submit(piece, base)


class _Slot:
    """Pinned host staging of one pipeline piece (bases + offsets in, offsets + int16 signal out).  Slots are recycled:
    in steady state NO pinned memory is allocated — cudaHostAlloc / cudaFreeHost synchronise the whole device, and a
    miss in torch's pinned-memory cache in the middle of a run cost the host its lead over the GPU (measured: e2e
    runs 25 % slower at random)."""

    def __init__(self):
        self.bases = self.ro = self.co = self.off = self.sig = None

    @staticmethod
    def _grown(t, n, dtype, quantum):
        if t is not None and t.numel() >= n:
            return t, False
        cap = -(-int(1.25 * n + 1) // quantum) * quantum
        return torch.empty(cap, dtype=dtype, pin_memory=True), True

    def fit_inputs(self, n_bases, n_reads):
        self.bases, a = self._grown(self.bases, max(n_bases, 1), torch.uint8, 1 << 20)
        self.ro, b = self._grown(self.ro, n_reads + 1, torch.int64, 1 << 12)
        self.co, c = self._grown(self.co, n_reads + 1, torch.int64, 1 << 12)
        self.off, d = self._grown(self.off, n_reads + 1, torch.int64, 1 << 12)
        return a + b + c + d

    def fit_signal(self, n):
        self.sig, a = self._grown(self.sig, max(n, 1), torch.int16, 8 << 20)
        return int(a)


class _ReadPipeline:
    """compute stream: H2D(bases) -> s2s_forward_reads;  copy stream: D2H(offsets) -> D2H(int16 prefix);  writer
    thread: writer.signals = {...}; writer.save().  PIPE_DEPTH pieces of compute are queued ahead of the copies."""

    def __init__(self, model: seq2squiggle):
        self.m = model
        self.eng = model.engine
        self.dev = model.device
        self.compute = torch.cuda.Stream(self.dev)
        self.copy = torch.cuda.Stream(self.dev)
        self.inflight: list = []     # pieces whose compute is queued but whose signal has not been fetched
        self.q: "queue.Queue" = queue.Queue(maxsize=4)
        self.err: Optional[BaseException] = None
        self.stats = dict(reads=0, chunks=0, samples=0, h2d_bytes=0, d2h_bytes=0, pinned_allocs=0)
        # the slot pool belongs to the model, so it survives the pipeline object of one predict epoch
        if not hasattr(model, "_slot_pool"):
            model._slot_pool = queue.Queue()
        self.free_slots: "queue.Queue" = model._slot_pool
        self.thread = threading.Thread(target=self._writer_loop, daemon=True)
        self.thread.start()

    def _slot(self) -> _Slot:
        try:
            return self.free_slots.get_nowait()
        except queue.Empty:
            return _Slot()

    def submit(self, reads, chunk_id_base=None):
        if self.err:
            raise self.err
        m = self.m
        names = [n for _, n in reads]
        joined, read_off, chunk_off = Engine.pack_reads_np([s for s, _ in reads], self.eng.k)
        n_reads, n_chunks, n_bases = len(names), int(chunk_off[-1]), len(joined)
        slot = self._slot()
        self.stats["pinned_allocs"] += slot.fit_inputs(n_bases, n_reads)
        slot.bases[:n_bases].numpy()[:] = np.frombuffer(joined, dtype=np.uint8)
        slot.ro[:n_reads + 1].numpy()[:] = read_off
        slot.co[:n_reads + 1].numpy()[:] = chunk_off
        base = m.chunks_done if chunk_id_base is None else chunk_id_base
        m.chunks_done = base + n_chunks
        with torch.cuda.stream(self.compute):
            d = [t.to(self.dev, non_blocking=True) for t in (slot.bases[:max(n_bases, 1)], slot.ro[:n_reads + 1],
                                                              slot.co[:n_reads + 1])]
            raw, raw_off, _ = self.eng.forward_reads_device(d[0], d[1], d[2], n_reads, n_chunks, m.run_options(), base)
            done = torch.cuda.Event()
            done.record(self.compute)
        off_host = slot.off[:n_reads + 1]
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(done)
            off_host.copy_(raw_off, non_blocking=True)
            off_ev = torch.cuda.Event()
            off_ev.record(self.copy)
        cur = dict(names=names, raw=raw, raw_off=raw_off, off_host=off_host, off_ev=off_ev, keep=d, slot=slot,
                   n_chunks=n_chunks)
        self.stats["h2d_bytes"] += n_bases + 16 * (n_reads + 1)
        self.inflight.append(cur)
        while len(self.inflight) > PIPE_DEPTH:      # the host runs PIPE_DEPTH pieces ahead of the device
            self._fetch(self.inflight.pop(0))

    def _fetch(self, b):
        b["off_ev"].synchronize()
        n = int(b["off_host"][-1])
        self.stats["pinned_allocs"] += b["slot"].fit_signal(n)
        sig = b["slot"].sig
        with torch.cuda.stream(self.copy):
            sig[:n].copy_(b["raw"][:n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy)
        b.update(sig=sig, n=n, sig_ev=ev)
        self.stats["d2h_bytes"] += 2 * n + 8 * b["off_host"].numel()
        self.stats["reads"] += len(b["names"]); self.stats["chunks"] += b["n_chunks"]; self.stats["samples"] += n
        self.q.put(b)

    def _writer_loop(self):
        while True:
            b = self.q.get()
            if b is None:
                return
            try:
                b["sig_ev"].synchronize()
                b["raw"] = b["raw_off"] = b["keep"] = None            # device buffers back to the allocator
                off = b["off_host"].numpy()
                sig = b["sig"].numpy()
                w = self.m.out_writer
                if w is not None:      # the views are valid until save() returns (writer plug-point contract)
                    w.signals = OrderedDict((name, sig[off[i]:off[i + 1]]) for i, name in enumerate(b["names"]))
                    w.save()
                    w.signals = []
                b.pop("sig")
                self.free_slots.put(b.pop("slot"))
            except BaseException as exc:  # surfaced on the next submit()/finish()
                self.err = exc

    def finish(self):
        while self.inflight:
            self._fetch(self.inflight.pop(0))
        self.q.put(None)
        self.thread.join()
        self.eng.check()
        if self.err:
            raise self.err
