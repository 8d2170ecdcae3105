"""ctypes binding of ``libs2s_b200.so`` (the C-ABI in ``include/s2s_b200.h``) and its nvcc build recipe."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# S2S_LIB_PATH: developer override to build / load an experimental variant next to the shipped library
LIB_PATH = os.environ.get("S2S_LIB_PATH") or os.path.join(HERE, "libs2s_b200.so")
SOURCES = ["s2s_api.cu", "k_frontend.cu", "k_simt.cu", "k_samplers.cu", "k_length_regulate.cu", "k_epilogue.cu",
           "k_tc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared"]

EXPORTS = ["s2s_last_error", "s2s_abi_version", "s2s_weights_count", "s2s_create", "s2s_destroy",
           "s2s_workspace_bytes", "s2s_chunks_of_read", "s2s_forward_reads", "s2s_forward_chunks",
           "s2s_check", "s2s_length_regulate", "s2s_digitise", "s2s_compact_reads", "s2s_profile_kernel", "s2s_profile_kernel_group", "s2s_launch_count", "s2s_debug_counters"]


class S2SConfig(C.Structure):
    _fields_ = [("seq_kmer", C.c_int32), ("encoder_layers", C.c_int32), ("decoder_layers", C.c_int32),
                ("pre_layers", C.c_int32), ("dmodel", C.c_int32), ("dff", C.c_int32), ("heads", C.c_int32),
                ("max_dna_len", C.c_int32), ("max_signal_len", C.c_int32), ("scaling_max_value", C.c_float)]


class S2SRunOpts(C.Structure):
    _fields_ = [("duration_mode", C.c_int32), ("dwell_mean", C.c_float), ("dwell_std", C.c_float),
                ("min_duration", C.c_float), ("noise_mode", C.c_int32), ("noise_std", C.c_float),
                ("min_noise", C.c_float), ("digitisation", C.c_float), ("range", C.c_float),
                ("offset_mean", C.c_float), ("rna_reverse", C.c_int32), ("seed", C.c_uint64),
                ("chunk_id_base", C.c_uint64), ("precision", C.c_int32)]


TAP_FIELDS = ["emb_out", "enc_out", "sigma", "conc", "rate", "dur_float", "dur_int", "lr_out", "sigma_ext", "p", "pa"]


class S2STaps(C.Structure):
    _fields_ = [(name + "_dev", C.c_void_p) for name in TAP_FIELDS]


BLOW5_LIB_PATH = os.path.join(HERE, "libs2s_blow5.so")
BLOW5_SOURCE = "blow5_writer.cpp"
EXPORTS_BLOW5 = ["s2s_blow5_last_error", "s2s_blow5_open", "s2s_blow5_write_batch", "s2s_blow5_bytes_written",
                 "s2s_blow5_close", "s2s_blow5_header", "s2s_blow5_encode_batch", "s2s_blow5_free"]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f != BLOW5_SOURCE] + \
           [os.path.join(HERE, "..", "include", "s2s_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every kernel for sm_100a into the in-tree ``libs2s_b200.so`` (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    extra = os.environ.get("S2S_NVCC_EXTRA", "").split()
    cmd = ["nvcc"] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_PATH


def build_blow5(force: bool = False) -> str:
    """Compile the native SLOW5/BLOW5 writer (host C++, zlib) into the in-tree ``libs2s_blow5.so``."""
    src, hdr = os.path.join(CSRC, BLOW5_SOURCE), os.path.join(HERE, "..", "include", "s2s_blow5.h")
    if not force and os.path.exists(BLOW5_LIB_PATH) and \
            os.path.getmtime(BLOW5_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return BLOW5_LIB_PATH
    cmd = ["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", BLOW5_LIB_PATH, src, "-lz"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ (blow5_writer) failed:\n" + res.stdout + res.stderr)
    return BLOW5_LIB_PATH


_blow5 = None


def load_blow5() -> C.CDLL:
    global _blow5
    if _blow5 is not None:
        return _blow5
    if not os.path.exists(BLOW5_LIB_PATH):
        raise RuntimeError(f"{BLOW5_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(BLOW5_LIB_PATH)
    vp, i64, i32, f64 = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    lib.s2s_blow5_last_error.restype = C.c_char_p
    lib.s2s_blow5_open.restype = C.c_int
    lib.s2s_blow5_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(vp)]
    lib.s2s_blow5_write_batch.restype = C.c_int
    lib.s2s_blow5_write_batch.argtypes = [vp, i64, C.c_char_p, vp, vp, vp, vp, vp, vp, f64, f64, f64, i32]
    lib.s2s_blow5_header.restype = C.c_int
    lib.s2s_blow5_header.argtypes = [C.c_int, C.c_int, C.c_char_p, C.POINTER(vp), C.POINTER(i64)]
    lib.s2s_blow5_encode_batch.restype = C.c_int
    lib.s2s_blow5_encode_batch.argtypes = [C.c_int, C.c_int, i64, C.c_char_p, vp, vp, vp, vp, vp, vp, f64, f64, f64, i32,
                                           C.POINTER(vp), C.POINTER(i64)]
    lib.s2s_blow5_free.restype = None
    lib.s2s_blow5_free.argtypes = [vp]
    lib.s2s_blow5_bytes_written.restype = i64
    lib.s2s_blow5_bytes_written.argtypes = [vp]
    lib.s2s_blow5_close.restype = C.c_int
    lib.s2s_blow5_close.argtypes = [vp]
    _blow5 = lib
    return lib


def check_blow5(status: int, what: str) -> None:
    if status != 0:
        raise RuntimeError(f"{what} failed ({status}): {load_blow5().s2s_blow5_last_error().decode('utf-8', 'replace')}")


_lib = None


def load() -> C.CDLL:
    """Load the extension; a missing library is a hard error (no CPU / PyTorch fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(seq2squiggle_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float
    lib.s2s_last_error.restype = C.c_char_p
    lib.s2s_abi_version.restype = C.c_int
    lib.s2s_launch_count.restype = i64
    lib.s2s_weights_count.restype = i64
    lib.s2s_weights_count.argtypes = [C.POINTER(S2SConfig)]
    lib.s2s_create.restype = C.c_int
    lib.s2s_create.argtypes = [vp, i64, C.POINTER(S2SConfig), C.c_int, C.POINTER(vp)]
    lib.s2s_destroy.restype = None
    lib.s2s_destroy.argtypes = [vp]
    lib.s2s_workspace_bytes.restype = i64
    lib.s2s_workspace_bytes.argtypes = [vp, i64, i64]
    lib.s2s_chunks_of_read.restype = i64
    lib.s2s_chunks_of_read.argtypes = [i64, i32]
    lib.s2s_forward_reads.restype = C.c_int
    lib.s2s_forward_reads.argtypes = [vp, vp, vp, vp, i64, i64, C.POINTER(S2SRunOpts), vp, i64, vp, vp,
                                      C.POINTER(S2STaps), vp]
    lib.s2s_forward_chunks.restype = C.c_int
    lib.s2s_forward_chunks.argtypes = [vp, vp, i64, C.POINTER(S2SRunOpts), vp, i64, vp, C.POINTER(S2STaps), vp]
    lib.s2s_check.restype = C.c_int
    lib.s2s_check.argtypes = [vp, vp]
    lib.s2s_length_regulate.restype = C.c_int
    lib.s2s_length_regulate.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp]
    lib.s2s_digitise.restype = C.c_int
    lib.s2s_digitise.argtypes = [vp, i64, f32, f32, f32, vp, vp]
    lib.s2s_compact_reads.restype = C.c_int
    lib.s2s_compact_reads.argtypes = [vp, vp, i64, i64, f32, f32, f32, i32, vp, i64, vp, vp, vp]
    lib.s2s_profile_kernel.restype = C.c_int
    lib.s2s_profile_kernel.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(i64)]
    lib.s2s_debug_counters.restype = C.c_int
    lib.s2s_debug_counters.argtypes = [vp, i32, i32]
    if lib.s2s_abi_version() != 1:
        raise RuntimeError("libs2s_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def last_error() -> str:
    return load().s2s_last_error().decode("utf-8", "replace")


def check(status: int, what: str) -> None:
    if status != 0:
        raise RuntimeError(f"{what} failed ({status}): {last_error()}")
