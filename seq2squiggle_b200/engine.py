"""Host-side driver of the C-ABI: owns the handle, the device workspace and the run options.

PyTorch is used only for device memory, pinned staging buffers and streams (``tensor.data_ptr()`` /
``torch.cuda.current_stream().cuda_stream`` are what crosses the C-ABI).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .checkpoint import check_architecture, pack_weights

DUR_CONSTANT, DUR_NORMAL, DUR_SAMPLER = 0, 1, 2
NOISE_OFF, NOISE_STATIC, NOISE_SAMPLER = 0, 1, 2
PREC_FP16_TC, PREC_FP32 = 0, 1
PRECISIONS = {"fp16": PREC_FP16_TC, "fp16-tc": PREC_FP16_TC, "fp32": PREC_FP32}

TAP_SHAPES = {  # per chunk
    "emb_out": ((16, 64), torch.float32), "enc_out": ((16, 64), torch.float32), "sigma": ((16,), torch.float32),
    "conc": ((16,), torch.float32), "rate": ((16,), torch.float32), "dur_float": ((16,), torch.float32),
    "dur_int": ((16,), torch.int32), "lr_out": ((250, 64), torch.float32), "sigma_ext": ((250,), torch.float32),
    "p": ((250,), torch.float32), "pa": ((250,), torch.float32),
}


@dataclass
class RunOptions:
    """The predict options that parameterise the hot path (seq2squiggle.py:230-390 / inference.py:348-368)."""
    dwell_mean: float
    dwell_std: float = 0.0
    duration_sampling: bool = True
    min_duration: float = 3
    noise_std: float = 2.0
    noise_sampling: bool = True
    min_noise: float = 0.0
    digitisation: float = 2048.0
    range: float = 281.345551
    offset_mean: float = -127.5655735
    rna: bool = False
    seed: int = 1
    precision: str = "fp16"

    @classmethod
    def from_profile(cls, profile: dict, profile_name: str, *, dwell_mean=None, **kw) -> "RunOptions":
        if dwell_mean is None:
            dwell_mean = profile["sample_rate"] / profile["bps"]  # inference.py:358-359
        return cls(dwell_mean=float(dwell_mean), digitisation=float(profile["digitisation"]),
                   range=float(profile["range"]), offset_mean=float(profile["offset_mean"]),
                   rna=profile_name.startswith("rna"), **kw)

    def to_c(self, chunk_id_base: int = 0) -> _lib.S2SRunOpts:
        if self.duration_sampling:                      # modules.py:410
            dmode = DUR_SAMPLER
        elif self.dwell_std <= 0:                       # modules.py:419
            dmode = DUR_CONSTANT
        else:
            dmode = DUR_NORMAL
        if not (self.noise_std > 0):                    # model.py:224
            nmode = NOISE_OFF
        else:
            nmode = NOISE_SAMPLER if self.noise_sampling else NOISE_STATIC
        if self.precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        return _lib.S2SRunOpts(dmode, self.dwell_mean, self.dwell_std, float(self.min_duration), nmode,
                               self.noise_std, self.min_noise, self.digitisation, self.range, self.offset_mean,
                               int(self.rna), int(self.seed) & 0xFFFFFFFFFFFFFFFF, int(chunk_id_base),
                               PRECISIONS[self.precision])


def chunks_of_read(read_len: int, k: int) -> int:
    n = read_len - k + 1
    return 0 if n <= 0 else -(-n // 16)


class Engine:
    """One engine per GPU (= per process).  Not thread-safe."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], config: dict, device: int = 0):
        check_architecture(config)
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("seq2squiggle_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        self.config = dict(config)
        self.k = int(config["seq_kmer"])
        self.cfg_c = _lib.S2SConfig(self.k, int(config["encoder_layers"]), int(config["decoder_layers"]),
                                    int(config["pre_layers"]), 64, 256, 8, 16, 250, float(config["scaling_max_value"]))
        blob = pack_weights(state_dict, config)
        expect = self.lib.s2s_weights_count(C.byref(self.cfg_c))
        if expect != blob.size:
            raise RuntimeError(f"weight blob size {blob.size} != {expect}")
        handle = C.c_void_p()
        torch.cuda.set_device(self.device)
        _lib.check(self.lib.s2s_create(blob.ctypes.data_as(C.c_void_p), blob.size, C.byref(self.cfg_c), device,
                                       C.byref(handle)), "s2s_create")
        self.handle = handle
        self._ws: Optional[torch.Tensor] = None

    def close(self):
        if getattr(self, "handle", None):
            self.lib.s2s_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def _workspace(self, n_chunks: int, n_reads: int) -> torch.Tensor:
        need = self.lib.s2s_workspace_bytes(self.handle, n_chunks, n_reads)
        if need < 0:
            raise RuntimeError("s2s_workspace_bytes failed")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            # grow geometrically: a reallocation is a cudaFree + cudaMalloc, i.e. a device synchronisation
            self._ws = torch.empty(int(need * 1.3) + 4096, dtype=torch.uint8, device=self.device)
        return self._ws

    def _make_taps(self, n_chunks: int, names) -> Tuple[Optional[_lib.S2STaps], Dict[str, torch.Tensor]]:
        if not names:
            return None, {}
        names = list(TAP_SHAPES) if names is True else list(names)
        out, taps = {}, _lib.S2STaps()
        for name in names:
            shape, dt = TAP_SHAPES[name]
            t = torch.zeros((n_chunks,) + shape, dtype=dt, device=self.device)
            out[name] = t
            setattr(taps, name + "_dev", t.data_ptr())
        return taps, out

    @staticmethod
    def pack_reads_np(reads: Sequence, k: int):
        """pack_reads without torch: (joined bytes, read offsets int64 [n+1], chunk offsets int64 [n+1])."""
        n = len(reads)
        if n and all(type(r) is str for r in reads):
            lens = np.fromiter(map(len, reads), dtype=np.int64, count=n)
            joined = "".join(reads).encode("latin-1", "replace")
        else:
            bufs = [r.encode("latin-1", "replace") if isinstance(r, str) else bytes(r) for r in reads]
            lens = np.fromiter((len(b) for b in bufs), dtype=np.int64, count=n)
            joined = b"".join(bufs)
        read_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=read_off[1:])
        nk = lens - k + 1
        nch = np.where(nk > 0, (nk + 15) // 16, 0)
        chunk_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(nch, out=chunk_off[1:])
        return joined, read_off, chunk_off

    @staticmethod
    def pack_reads(reads: Sequence, k: int, pin: bool = False):
        """Host side of the boundary: concatenate read bytes, prefix offsets of bases and of chunks."""
        n = len(reads)
        if n and all(type(r) is str for r in reads):
            # one join + one encode (latin-1 is 1 byte per character, so lengths are the string lengths)
            lens = np.fromiter(map(len, reads), dtype=np.int64, count=n)
            bufs = ["".join(reads).encode("latin-1", "replace")]
        else:
            bufs = [r.encode("latin-1", "replace") if isinstance(r, str) else bytes(r) for r in reads]
            lens = np.fromiter((len(b) for b in bufs), dtype=np.int64, count=n)
        read_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=read_off[1:])
        nk = lens - k + 1
        nch = np.where(nk > 0, (nk + 15) // 16, 0)
        chunk_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(nch, out=chunk_off[1:])
        bases = torch.frombuffer(bytearray(b"".join(bufs)) or bytearray(1), dtype=torch.uint8)
        ro, co = torch.from_numpy(read_off), torch.from_numpy(chunk_off)
        if pin:
            bases, ro, co = bases.pin_memory(), ro.pin_memory(), co.pin_memory()
        return bases, ro, co

    def check(self):
        """Synchronise and fail loudly if a kernel raised its device-side error word (bounded barrier wait)."""
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.s2s_check(self.handle, st), "s2s_check")

    def forward_reads_device(self, bases: torch.Tensor, read_off: torch.Tensor, chunk_off: torch.Tensor,
                             n_reads: int, n_chunks: int, opts: RunOptions, chunk_id_base: int = 0, taps=None, out=None):
        """All inputs already on the device.  Returns (raw int16 [n_chunks*250 cap], raw_offsets int64 [n_reads+1],
        taps dict); the valid prefix of raw is raw_offsets[-1] samples.  ``out=(raw, raw_offsets)``: caller-owned
        output buffers (at least n_chunks*250 / n_reads+1 elements) instead of fresh allocations."""
        ws = self._workspace(n_chunks, n_reads)
        if out is not None:
            raw, raw_off = out
            assert raw.dtype == torch.int16 and raw.numel() >= n_chunks * 250 and raw_off.numel() >= n_reads + 1
        else:
            raw = torch.empty(max(n_chunks * 250, 1), dtype=torch.int16, device=self.device)
            raw_off = torch.empty(n_reads + 1, dtype=torch.int64, device=self.device)
        taps_c, tap_out = self._make_taps(n_chunks, taps)
        o = opts.to_c(chunk_id_base)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.s2s_forward_reads(self.handle, bases.data_ptr(), read_off.data_ptr(), chunk_off.data_ptr(),
                                              n_reads, n_chunks, C.byref(o), ws.data_ptr(), ws.numel(), raw.data_ptr(),
                                              raw_off.data_ptr(), C.byref(taps_c) if taps_c else None, st),
                   "s2s_forward_reads")
        return raw, raw_off, tap_out

    def forward_reads(self, reads: Sequence, opts: RunOptions, chunk_id_base: int = 0, taps=None):
        """Host reads in, host signals out: list of int16 numpy arrays, one per read (empty for skipped reads)."""
        bases, ro, co = self.pack_reads(reads, self.k, pin=True)
        n_reads, n_chunks = len(reads), int(co[-1])
        raw, raw_off, tap_out = self.forward_reads_device(
            bases.to(self.device, non_blocking=True), ro.to(self.device, non_blocking=True),
            co.to(self.device, non_blocking=True), n_reads, n_chunks, opts, chunk_id_base, taps)
        self.check()
        off = raw_off.cpu().numpy()
        sig = raw[: int(off[-1])].cpu().numpy()
        return [sig[off[i]:off[i + 1]] for i in range(n_reads)], tap_out

    def forward_chunks(self, codes: torch.Tensor, opts: RunOptions, chunk_id_base: int = 0, taps=None,
                       check: bool = True):
        """codes: int8 [C,16,k] on the device (argmax of the one-hot, -1 = zero row) -> pA float32 [C,250]."""
        assert codes.dtype == torch.int8 and codes.is_cuda and codes.is_contiguous()
        n_chunks = codes.shape[0]
        assert codes.shape[1:] == (16, self.k), f"expected [C,16,{self.k}] codes, got {tuple(codes.shape)}"
        ws = self._workspace(n_chunks, 0)
        pa = torch.empty((n_chunks, 250), dtype=torch.float32, device=self.device)
        taps_c, tap_out = self._make_taps(n_chunks, taps)
        o = opts.to_c(chunk_id_base)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.s2s_forward_chunks(self.handle, codes.data_ptr(), n_chunks, C.byref(o), ws.data_ptr(),
                                               ws.numel(), pa.data_ptr(), C.byref(taps_c) if taps_c else None, st),
                   "s2s_forward_chunks")
        if check:
            self.check()
        return pa, tap_out

    # ---- stage entry points ------------------------------------------------------------------
    def length_regulate(self, x: torch.Tensor, sigma: torch.Tensor, dur: torch.Tensor):
        n = x.shape[0]
        out = torch.empty((n, 250, 64), dtype=torch.float32, device=self.device)
        sext = torch.empty((n, 250), dtype=torch.float32, device=self.device)
        total = torch.empty((n,), dtype=torch.int32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.s2s_length_regulate(x.contiguous().data_ptr(), sigma.contiguous().data_ptr(),
                                                dur.contiguous().data_ptr(), n, out.data_ptr(), sext.data_ptr(),
                                                total.data_ptr(), st), "s2s_length_regulate")
        return out, sext, total

    def digitise(self, pa: torch.Tensor, digitisation: float, signal_range: float, offset: float) -> torch.Tensor:
        pa = pa.contiguous().to(torch.float32)
        raw = torch.empty(pa.shape, dtype=torch.int16, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.s2s_digitise(pa.data_ptr(), pa.numel(), digitisation, signal_range, offset,
                                         raw.data_ptr(), st), "s2s_digitise")
        return raw

    def compact_reads(self, pa: torch.Tensor, chunk_off: torch.Tensor, opts: RunOptions):
        n_chunks, n_reads = pa.shape[0], chunk_off.numel() - 1
        ws = self._workspace(n_chunks, n_reads)
        raw = torch.empty(max(n_chunks * 250, 1), dtype=torch.int16, device=self.device)
        raw_off = torch.empty(n_reads + 1, dtype=torch.int64, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.s2s_compact_reads(pa.contiguous().data_ptr(), chunk_off.data_ptr(), n_reads, n_chunks,
                                              opts.digitisation, opts.range, opts.offset_mean, int(opts.rna),
                                              ws.data_ptr(), ws.numel(), raw.data_ptr(), raw_off.data_ptr(), st),
                   "s2s_compact_reads")
        return raw, raw_off

    def launch_count(self) -> int:
        return int(self.lib.s2s_launch_count())
