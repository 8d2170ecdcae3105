"""SLOW5 / BLOW5 / POD5 writers with the reference's plug-point contract (``signal_io.py:62-282``).

``writer.signals = {read_id: 1-D signal}`` then ``writer.save()`` — exactly how ``export_and_clear_results``
(``model.py:290-292``) drives the reference writers.  Two kinds of signal values are accepted:

* **int16** (numpy array or tensor): already zero-stripped, digitised (and reversed for RNA) by the fused CUDA
  kernel behind ``s2s_forward_reads`` / ``s2s_compact_reads`` — the normal case in this package;
* **float** pA tensors on the device (the reference contract): digitised here through the ``s2s_digitise`` CUDA
  kernel.  There is no NumPy/CPU digitisation path: without the CUDA library this raises.

The container bytes are produced by the native writer in ``csrc/blow5_writer.cpp`` (``include/s2s_blow5.h``)
instead of pyslow5; record fields follow ``signal_io.py:104-171``.  POD5 needs the third-party ``pod5`` package
(Arrow based); it is used when importable and refused with a clear error otherwise.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import uuid
from datetime import datetime
from typing import Optional

import numpy as np

from . import _lib

logger = logging.getLogger("seq2squiggle")


def indexed_uuid(index: int) -> uuid.UUID:
    """signal_io.py:19-23."""
    return uuid.UUID(f"00000000-0000-0000-0000-{index:012d}")


_KITS = {  # signal_io.py:30-51
    "rna-004": ("sqk-rna004", "FLO-PRO004RA", "FLO-MIN004RA"),
    "rna-002": ("sqk-rna002", "FLO-PRO002", "FLO-MIN106"),
    "dna-r10": ("SQK-LSK114", "FLO-PRO114", "FLO-MIN114"),
    "dna-r9": ("SQK-LSK109", "FLO-PRO001", "FLO-MIN110"),
}


def get_seq_kit_and_flow_cell(profile_name: str):
    """signal_io.py:26-60."""
    for prefix, (kit, prom, minion) in _KITS.items():
        if profile_name.startswith(prefix):
            if "prom" in profile_name:
                return kit, prom
            if "min" in profile_name:
                return kit, minion
            break
    raise ValueError(f"Unsupported profile name: {profile_name}")


def _as_int16_or_device_float(signal):
    """Returns ('raw', int16 ndarray) or ('pa', float32 CUDA tensor)."""
    import torch
    if isinstance(signal, np.ndarray):
        if signal.dtype != np.int16:
            raise TypeError("host signals must already be int16 (digitised on the GPU); float pA must be a CUDA tensor")
        return "raw", signal.reshape(-1)
    if isinstance(signal, torch.Tensor):
        if signal.dtype == torch.int16:
            return "raw", signal.detach().reshape(-1).cpu().numpy()
        if not signal.is_cuda:
            raise RuntimeError("float pA signals must live on the CUDA device: seq2squiggle_b200 digitises with its "
                               "CUDA kernel and has no CPU fallback")
        return "pa", signal.detach().reshape(-1).to(torch.float32)
    raise TypeError(f"unsupported signal type {type(signal)!r}")


class _WriterBase:
    def __init__(self, filename, profile, ideal_mode, profile_name, preserve_read_ids):
        self.filename = filename
        self.profile: dict = profile
        self.ideal_mode = ideal_mode
        self.profile_name = profile_name
        self.preserve_read_ids = preserve_read_ids
        self.signals = None
        self.median_before = float(profile["median_before_mean"])
        self.median_before_std = float(profile["median_before_std"])
        self.offset = float(profile["offset_mean"])
        self.offset_std = float(profile["offset_std"])
        self.digitisation = float(profile["digitisation"])
        self.signal_range = float(profile["range"])
        self.sample_rate = float(profile["sample_rate"])
        self.start_time = 0
        self.reads_written = 0
        self.samples_written = 0
        self.is_rna = profile_name.startswith("rna")
        # The reference restarts `idx` at every save() (signal_io.py:123), so a run with more than one flush
        # repeats read ids / read numbers.  Here numbering continues across flushes (identical for one flush).
        self._id_base = 0

    # signal_io.py:128-141: per-read raw int16 in `signals` order; float pA goes through the CUDA digitiser
    def _collect(self):
        import torch
        kept = []  # (idx, read_id, kind, payload)
        base = self._id_base
        self._id_base += len(self.signals)
        for idx, (read_id, signal) in enumerate(self.signals.items(), start=base):
            kind, payload = _as_int16_or_device_float(signal)
            if payload.shape[0] == 0:
                logger.debug("Empty signal, skipping {}".format(read_id))
                continue
            kept.append((idx, read_id, kind, payload))
        pa_items = [k for k in kept if k[2] == "pa"]
        if pa_items:
            lib = _lib.load()
            cat = torch.cat([k[3] for k in pa_items]).contiguous()
            raw = torch.empty(cat.shape, dtype=torch.int16, device=cat.device)
            st = torch.cuda.current_stream(cat.device).cuda_stream
            _lib.check(lib.s2s_digitise(cat.data_ptr(), cat.numel(), self.digitisation, self.signal_range, self.offset,
                                        raw.data_ptr(), st), "s2s_digitise")
            host = raw.cpu().numpy()
            pos, out = 0, {}
            for idx, _, _, payload in pa_items:
                n = payload.shape[0]
                sig = host[pos:pos + n]
                out[idx] = np.ascontiguousarray(sig[::-1]) if self.is_rna else sig
                pos += n
            kept = [(i, r, "raw", out[i] if k == "pa" else p) for i, r, k, p in kept]
        return [(i, r, p) for i, r, _, p in kept]

    def _read_meta(self):
        if self.ideal_mode:
            return self.median_before, self.offset
        # signal_io.py:131-133: median_before first, then offset (global NumPy generator seeded by set_seeds)
        return (np.random.normal(self.median_before, self.median_before_std),
                np.random.normal(self.offset, self.offset_std))


class SharedOrder:
    """Hand-over table of a multi-process run that writes ONE output file: for every batch ``b`` (batches are dealt to
    the ranks round-robin, in read order) the totals of all batches before it — emitted samples (the records'
    ``start_time``), written records (position in the per-record NumPy draw stream) and the byte offset in the file.
    The owner of batch ``b`` publishes the first two for ``b + 1`` as soon as its signal is on the host and the third as
    soon as its records are encoded, so the ranks encode and ``pwrite`` in parallel and nothing is spliced afterwards.
    The table is an ``int64`` memory map shared by the ranks of one node (x86 stores are ordered: a row's flag is written
    after its values)."""
    COLS = 6   # samples, records, ready1 | offset, ready2, (spare)

    def __init__(self, path: str, n_batches: int, create: bool):
        import numpy as np
        self.path, self.n_batches = path, int(n_batches)
        shape = (self.n_batches + 1, self.COLS)
        if create:
            with open(path, "wb") as f:
                f.truncate(8 * shape[0] * shape[1])
        self.tab = np.memmap(path, dtype=np.int64, mode="r+", shape=shape)

    def _wait(self, b: int, col: int, timeout: float = 3600.0):
        import time
        t0 = time.monotonic()
        while not self.tab[b, col]:
            if time.monotonic() - t0 > timeout:
                raise TimeoutError(f"ordered write: batch {b} was never released (did another rank fail?)")
            time.sleep(2e-5)

    def start(self, header_bytes: int):
        self.tab[0, 0], self.tab[0, 1], self.tab[0, 3] = 0, 0, int(header_bytes)
        self.tab[0, 2] = self.tab[0, 4] = 1

    def totals_before(self, b: int):
        self._wait(b, 2)
        return int(self.tab[b, 0]), int(self.tab[b, 1])

    def publish_totals(self, b: int, samples: int, records: int):
        self.tab[b, 0], self.tab[b, 1] = int(samples), int(records)
        self.tab[b, 2] = 1

    def offset_of(self, b: int) -> int:
        self._wait(b, 4)
        return int(self.tab[b, 3])

    def publish_offset(self, b: int, offset: int):
        self.tab[b, 3] = int(offset)
        self.tab[b, 4] = 1

    def close(self):
        if self.tab is not None:
            self.tab.flush()
            del self.tab
            self.tab = None


class BLOW5Writer(_WriterBase):
    """Export signal predictions to a slow5/blow5 file (signal_io.py:62-172)."""

    def __init__(self, filename, profile, ideal_mode, profile_name, preserve_read_ids, n_threads: Optional[int] = None,
                 record_compression: Optional[str] = None):
        super().__init__(filename, profile, ideal_mode, profile_name, preserve_read_ids)
        self.n_threads = n_threads or min(os.cpu_count() or 1, 32)
        # record | signal compression.  pyslow5's defaults (what the reference's writer gets, signal_io.py:98-102) are
        # zlib records + svb-zd signal = "zlib+svb-zd"; "none" (the default here) writes raw records at memory speed
        comp = (record_compression or os.environ.get("S2S_BLOW5_COMPRESS", "none")).lower().replace("_", "-")
        table = {"none": 0, "zlib": 1, "svb-zd": 1 << 8, "zlib+svb-zd": 1 | (1 << 8), "svb-zd+zlib": 1 | (1 << 8),
                 "pyslow5": 1 | (1 << 8)}
        if comp not in table:
            raise ValueError("BLOW5 compression must be one of 'none', 'zlib', 'svb-zd', 'zlib+svb-zd'")
        self.record_compression = table[comp]
        self.shared: Optional[SharedOrder] = None   # multi-process run: ordered writes into one shared file
        self._fd = -1
        self._draws_used = 0                          # records whose (median_before, offset) draws this process has made

    def _header_attrs(self) -> str:
        seq_kit, flow_cell = get_seq_kit_and_flow_cell(self.profile_name)
        attrs = {  # signal_io.py:104-113
            "asic_id": "asic_id_0",
            "exp_start_time": datetime.now().strftime("%Y-%m-%dT%H:%M:%SZ"),
            "run_id": "run_id_0",
            "flow_cell_id": "FAN00000",
            "flow_cell_product_code": flow_cell,
            "experiment_type": "rna" if self.is_rna else "genomic_dna",
            "sample_frequency": int(self.sample_rate),
            "sequencing_kit": seq_kit,
        }
        return "".join(f"{k}\t{v}\n" for k, v in attrs.items() if v is not None)

    def save(self):
        if self.signals is None:
            logger.warning("SLOW5 was not exported. No signals were found")
            raise ValueError("SLOW5 was not exported. No signals were found")
        lib = _lib.load_blow5()
        filename = str(self.filename)
        append = os.path.exists(filename)
        logger.debug(f"File mode for saving: {'a' if append else 'w'}")
        fmt = 1 if filename.endswith(".slow5") else 0
        handle = C.c_void_p()
        _lib.check_blow5(lib.s2s_blow5_open(filename.encode(), fmt, int(append), self.record_compression,
                                            self._header_attrs().encode(), C.byref(handle)), "s2s_blow5_open")
        try:
            items = self._collect() if self.signals else []
            n = len(items)
            if n:
                ids = bytearray()
                offsets = np.zeros(n + 1, dtype=np.int64)
                off_v, med_v = np.empty(n, np.float64), np.empty(n, np.float64)
                rnum, stime = np.empty(n, np.int32), np.empty(n, np.uint64)
                for j, (idx, read_id, sig) in enumerate(items):
                    med_v[j], off_v[j] = self._read_meta()
                    rid = read_id if self.preserve_read_ids else indexed_uuid(idx + 1)
                    ids += str(rid).encode() + b"\0"
                    offsets[j + 1] = offsets[j] + sig.shape[0]
                    rnum[j] = idx
                    stime[j] = self.start_time
                    self.start_time += int(sig.shape[0])
                flat = np.ascontiguousarray(np.concatenate([s for _, _, s in items]), dtype=np.int16)
                p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
                _lib.check_blow5(lib.s2s_blow5_write_batch(handle, n, bytes(ids), p(flat), p(offsets), p(off_v), p(med_v),
                                                           p(rnum), p(stime), self.digitisation, self.signal_range,
                                                           self.sample_rate, self.n_threads), "s2s_blow5_write_batch")
                self.reads_written += n
                self.samples_written += int(offsets[-1])
        finally:
            _lib.check_blow5(lib.s2s_blow5_close(handle), "s2s_blow5_close")


    # ---- multi-process ordered writing ----------------------------------------------------------------------
    def begin_shared(self, shared: SharedOrder, rank: int):
        """Open the (one) output file for ordered writing; rank 0 writes the header and releases batch 0."""
        lib = _lib.load_blow5()
        self.shared = shared
        filename = str(self.filename)
        self._fmt = 1 if filename.endswith(".slow5") else 0
        if rank == 0:
            buf, size = C.c_void_p(), C.c_int64()
            _lib.check_blow5(lib.s2s_blow5_header(self._fmt, self.record_compression, self._header_attrs().encode(),
                                                  C.byref(buf), C.byref(size)), "s2s_blow5_header")
            try:
                with open(filename, "wb") as f:
                    f.write(C.string_at(buf, size.value))
                    if shared.n_batches == 0 and self._fmt == 0:
                        f.write(b"5WOLB")
            finally:
                lib.s2s_blow5_free(buf)
            shared.start(size.value)
        else:
            shared.offset_of(0)                 # the file exists once batch 0 has been released
        self._fd = os.open(filename, os.O_RDWR)

    def end_shared(self):
        if self._fd >= 0:
            os.close(self._fd)
            self._fd = -1
        if self.shared is not None:
            self.shared.close()

    def _save_flat_shared(self, names, flat, offsets, tag):
        """Batch ``b`` (reads ``read_base ..``) of a shared-file run: numbering by global read index, ``start_time`` and
        the position in the NumPy draw stream from the totals of the batches before it, bytes at the offset handed over
        by the previous batch's owner.  The file equals the one a single process writes for the same seed."""
        b, read_base = tag
        sh = self.shared
        n_all = len(names)
        lens = np.diff(np.asarray(offsets[: n_all + 1], dtype=np.int64))
        keep = np.flatnonzero(lens > 0)
        n = int(keep.size)
        klens = lens[keep]
        out_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(klens, out=out_off[1:])
        samples0, records0 = sh.totals_before(b)
        sh.publish_totals(b + 1, samples0 + int(out_off[-1]), records0 + n)     # the next owner can go on at once
        lib = _lib.load_blow5()
        buf, size = C.c_void_p(), C.c_int64(0)
        if n:
            if n == n_all:
                sig = np.ascontiguousarray(flat[int(offsets[0]):int(offsets[n_all])], dtype=np.int16)
            else:
                sig = np.ascontiguousarray(np.concatenate([flat[int(offsets[i]):int(offsets[i + 1])] for i in keep]),
                                           dtype=np.int16)
            if self.ideal_mode:
                med_v = np.full(n, self.median_before, np.float64)
                off_v = np.full(n, self.offset, np.float64)
            else:   # the single-process stream: one (median_before, offset) pair per written record, in record order
                skip = records0 - self._draws_used
                if skip > 0:
                    np.random.standard_normal(2 * skip)
                draws = np.random.normal([self.median_before, self.offset], [self.median_before_std, self.offset_std],
                                         size=(n, 2))
                self._draws_used = records0 + n
                med_v, off_v = np.ascontiguousarray(draws[:, 0]), np.ascontiguousarray(draws[:, 1])
            idx = (read_base + keep).astype(np.int64)
            if self.preserve_read_ids:
                ids = b"".join(str(names[i]).encode() + b"\0" for i in keep)
            else:
                ids = b"".join(b"00000000-0000-0000-0000-%012d\0" % (int(i) + 1) for i in idx)
            rnum = idx.astype(np.int32)
            stime = (samples0 + out_off[:-1]).astype(np.uint64)
            p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
            _lib.check_blow5(lib.s2s_blow5_encode_batch(self._fmt, self.record_compression, n, ids, p(sig), p(out_off),
                                                        p(off_v), p(med_v), p(rnum), p(stime), self.digitisation,
                                                        self.signal_range, self.sample_rate, self.n_threads,
                                                        C.byref(buf), C.byref(size)), "s2s_blow5_encode_batch")
        try:
            at = sh.offset_of(b)
            sh.publish_offset(b + 1, at + size.value)
            if size.value and not os.environ.get("S2S_BLOW5_NULL_SINK"):    # (developer switch: encode, do not write)
                view = memoryview((C.c_char * size.value).from_address(buf.value))
                done = 0
                while done < size.value:
                    done += os.pwrite(self._fd, view[done:], at + done)
            if b == sh.n_batches - 1 and self._fmt == 0:
                os.pwrite(self._fd, b"5WOLB", at + size.value)
        finally:
            if buf.value:
                lib.s2s_blow5_free(buf)
        self.reads_written += n
        self.samples_written += int(out_off[-1])

    def save_flat(self, names, flat: np.ndarray, offsets: np.ndarray, tag=None):
        """Fast path of the read pipeline: the digitised signals of a batch as ONE contiguous int16 array plus per-read
        offsets (read i = flat[offsets[i]:offsets[i+1]], in `names` order).  Writes exactly the records that
        ``signals = {name: flat[...]}; save()`` writes — same numbering, same order of the per-record NumPy draws, empty
        reads skipped — with the per-read Python work (two scalar draws, a UUID object, a bytes append, a second copy
        of every signal) replaced by array operations: the writer thread shares the GIL with the read sampler."""
        if self.shared is not None:
            if tag is None:
                raise RuntimeError("shared-file writing needs the batch tag (batch index, first global read)")
            return self._save_flat_shared(names, flat, offsets, tag)
        n_all = len(names)
        lens = np.diff(np.asarray(offsets[: n_all + 1], dtype=np.int64))
        base = self._id_base
        self._id_base += n_all
        keep = np.flatnonzero(lens > 0)
        n = int(keep.size)
        lib = _lib.load_blow5()
        filename = str(self.filename)
        append = os.path.exists(filename)
        fmt = 1 if filename.endswith(".slow5") else 0
        handle = C.c_void_p()
        _lib.check_blow5(lib.s2s_blow5_open(filename.encode(), fmt, int(append), self.record_compression,
                                            self._header_attrs().encode(), C.byref(handle)), "s2s_blow5_open")
        try:
            if n:
                klens = lens[keep]
                if n == n_all:       # nothing skipped: the batch buffer is already the concatenation
                    sig = np.ascontiguousarray(flat[int(offsets[0]):int(offsets[n_all])], dtype=np.int16)
                else:
                    sig = np.ascontiguousarray(np.concatenate([flat[int(offsets[i]):int(offsets[i + 1])] for i in keep]),
                                               dtype=np.int16)
                out_off = np.zeros(n + 1, dtype=np.int64)
                np.cumsum(klens, out=out_off[1:])
                if self.ideal_mode:
                    med_v = np.full(n, self.median_before, np.float64)
                    off_v = np.full(n, self.offset, np.float64)
                else:   # (median_before, offset) per record, in record order: the same stream as n pairs of scalar draws
                    draws = np.random.normal([self.median_before, self.offset], [self.median_before_std, self.offset_std],
                                             size=(n, 2))
                    med_v, off_v = np.ascontiguousarray(draws[:, 0]), np.ascontiguousarray(draws[:, 1])
                idx = (base + keep).astype(np.int64)
                if self.preserve_read_ids:
                    ids = b"".join(str(names[i]).encode() + b"\0" for i in keep)
                else:
                    ids = b"".join(b"00000000-0000-0000-0000-%012d\0" % (int(i) + 1) for i in idx)
                rnum = idx.astype(np.int32)
                stime = (self.start_time + out_off[:-1]).astype(np.uint64)
                self.start_time += int(out_off[-1])
                p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
                _lib.check_blow5(lib.s2s_blow5_write_batch(handle, n, ids, p(sig), p(out_off), p(off_v), p(med_v),
                                                           p(rnum), p(stime), self.digitisation, self.signal_range,
                                                           self.sample_rate, self.n_threads), "s2s_blow5_write_batch")
                self.reads_written += n
                self.samples_written += int(out_off[-1])
        finally:
            _lib.check_blow5(lib.s2s_blow5_close(handle), "s2s_blow5_close")


class POD5Writer(_WriterBase):
    """Export signal predictions to a pod5 file (signal_io.py:175-282).  Needs the third-party ``pod5`` package."""
    appendable = False   # pod5.Writer refuses an existing file: the read pipeline keeps every read and saves once

    def save(self):
        if self.signals is None:
            logger.warning("POD5 was not exported. No signals were found")
            raise ValueError("POD5 was not exported. No signals were found")
        try:
            import pod5
        except ImportError as exc:  # pod5 is Arrow based and not part of this image
            raise RuntimeError("Writing .pod5 needs the 'pod5' package, which is not installed. Export to .blow5 and "
                               "convert with blue_crab instead.") from exc
        seq_kit, flow_cell = get_seq_kit_and_flow_cell(self.profile_name)
        now = datetime.now()
        run_info = pod5.RunInfo(
            acquisition_id="", acquisition_start_time=now, adc_max=4095, adc_min=-4096, context_tags={},
            experiment_name="", flow_cell_id="", flow_cell_product_code=flow_cell, protocol_name="",
            protocol_run_id="", protocol_start_time=now, sample_id="test", sample_rate=int(self.sample_rate),
            sequencing_kit=seq_kit, sequencer_position="", sequencer_position_type="", software="", system_name="",
            system_type="", tracking_id={})
        reads = []
        for idx, read_id, sig in self._collect():
            median_before, offset = self._read_meta()
            rid = uuid.uuid5(uuid.NAMESPACE_DNS, read_id) if self.preserve_read_ids else indexed_uuid(idx + 1)
            reads.append(pod5.Read(
                read_id=rid, pore=pod5.Pore(channel=123, well=3, pore_type="not_set"),
                calibration=pod5.Calibration(offset=offset, scale=self.signal_range / self.digitisation),
                read_number=idx, start_sample=0, median_before=median_before,
                end_reason=pod5.EndReason(reason=pod5.EndReasonEnum.SIGNAL_POSITIVE, forced=False),
                run_info=run_info, signal=sig))
        with pod5.Writer(self.filename) as writer:
            for read in reads:
                writer.add_read(read)
        self.reads_written += len(reads)
