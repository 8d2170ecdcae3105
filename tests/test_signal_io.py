"""SLOW5/BLOW5 writer (native libs2s_blow5.so behind seq2squiggle_b200.signal_io) read back with an independent
parser; record fields as the reference fills them (signal_io.py:104-171)."""

import os

import numpy as np
import pytest

from seq2squiggle_b200.profiles import get_profile
from seq2squiggle_b200.signal_io import BLOW5Writer, POD5Writer, get_seq_kit_and_flow_cell, indexed_uuid
from tests.blow5_reader import read_blow5, read_slow5


def _signals(rng, n, empty_at=()):
    out = {}
    for i in range(n):
        ln = 0 if i in empty_at else int(rng.integers(1, 700))
        out[f"read-{i}"] = rng.integers(-32768, 32767, size=ln).astype(np.int16)
    return out


@pytest.mark.parametrize("comp", ["none", "zlib", "svb-zd", "zlib+svb-zd"])
def test_blow5_roundtrip_ideal_mode(tmp_path, comp):
    rng = np.random.default_rng(0)
    prof = get_profile("dna-r10-prom")
    path = str(tmp_path / "o.blow5")
    w = BLOW5Writer(path, prof, True, "dna-r10-prom", False, n_threads=3, record_compression=comp)
    first = _signals(rng, 9, empty_at=(2,))
    w.signals = first
    w.save()
    second = _signals(rng, 5)
    second = {k + "b": v for k, v in second.items()}
    w.signals = second                       # second flush appends (file exists -> mode 'a')
    w.save()
    f = read_blow5(path)
    assert f["version"] == (0, 2, 0) and f["num_read_groups"] == 1
    assert f["signal_compression"] == (1 if "svb-zd" in comp else 0)        # pyslow5's default pair is zlib + svb-zd
    assert f["record_compression"] == (1 if "zlib" in comp else 0)
    assert f["attrs"]["asic_id"] == "asic_id_0" and f["attrs"]["run_id"] == "run_id_0"
    assert f["attrs"]["flow_cell_id"] == "FAN00000" and f["attrs"]["flow_cell_product_code"] == "FLO-PRO114"
    assert f["attrs"]["experiment_type"] == "genomic_dna" and f["attrs"]["sample_frequency"] == "5000"
    assert f["attrs"]["sequencing_kit"] == "SQK-LSK114" and "exp_start_time" in f["attrs"]
    assert f["names"] == ["read_id", "read_group", "digitisation", "offset", "range", "sampling_rate", "len_raw_signal",
                          "raw_signal", "channel_number", "median_before", "read_number", "start_mux", "start_time"]
    allsig = [(i, v) for i, v in enumerate(list(first.values()) + list(second.values()))]
    kept = [(i, v) for i, v in allsig if len(v)]
    assert len(f["records"]) == len(kept) == 13
    t = 0
    for rec, (idx, sig) in zip(f["records"], kept):
        assert rec["read_id"] == str(indexed_uuid(idx + 1))           # empty read 2 leaves a gap, like the reference
        assert rec["read_number"] == idx and rec["read_group"] == 0 and rec["start_mux"] == 0
        assert rec["channel_number"] == "0"
        assert rec["digitisation"] == 2048.0 and rec["range"] == prof["range"] and rec["sampling_rate"] == 5000.0
        assert rec["offset"] == prof["offset_mean"] and rec["median_before"] == prof["median_before_mean"]
        assert rec["signal"] == sig.tolist() and rec["len_raw_signal"] == len(sig)
        assert rec["start_time"] == t
        t += len(sig)
    assert w.samples_written == t and w.reads_written == 13


def test_blow5_non_ideal_metadata_draws_and_preserved_ids(tmp_path):
    prof = get_profile("rna-004-min")
    path = str(tmp_path / "o.blow5")
    w = BLOW5Writer(path, prof, False, "rna-004-min", True)
    sig = _signals(np.random.default_rng(1), 4)
    np.random.seed(42)
    w.signals = sig
    w.save()
    np.random.seed(42)
    exp = [(np.random.normal(prof["median_before_mean"], prof["median_before_std"]),
            np.random.normal(prof["offset_mean"], prof["offset_std"])) for _ in sig]   # signal_io.py:131-133 order
    f = read_blow5(path)
    assert f["attrs"]["experiment_type"] == "rna" and f["attrs"]["sequencing_kit"] == "sqk-rna004"
    assert f["attrs"]["flow_cell_product_code"] == "FLO-MIN004RA"
    for rec, name, (med, off) in zip(f["records"], sig, exp):
        assert rec["read_id"] == name and rec["median_before"] == med and rec["offset"] == off


def test_slow5_ascii(tmp_path):
    prof = get_profile("dna-r9-min")
    path = str(tmp_path / "o.slow5")
    w = BLOW5Writer(path, prof, True, "dna-r9-min", False)
    sig = _signals(np.random.default_rng(2), 3)
    w.signals = sig
    w.save()
    w.signals = {"x": np.array([1, -2, 3], dtype=np.int16)}
    w.save()
    f = read_slow5(path)
    assert f["attrs"]["sequencing_kit"] == "SQK-LSK109" and len(f["records"]) == 4
    for rec, s in zip(f["records"], list(sig.values()) + [np.array([1, -2, 3])]):
        assert rec["signal"] == s.tolist() and int(rec["len_raw_signal"]) == len(s)
        assert float(rec["range"]) == prof["range"] and float(rec["offset"]) == prof["offset_mean"]
    assert f["records"][3]["read_number"] == "3" and f["records"][3]["read_id"] == str(indexed_uuid(4))


def test_writer_errors_and_kits(tmp_path):
    prof = get_profile("dna-r10-min")
    w = BLOW5Writer(str(tmp_path / "e.blow5"), prof, True, "dna-r10-min", False)
    with pytest.raises(ValueError, match="No signals were found"):
        w.save()                                                       # signal_io.py:92-94
    with pytest.raises(ValueError, match="No signals were found"):
        POD5Writer(str(tmp_path / "e.pod5"), prof, True, "dna-r10-min", False).save()
    w.signals = {"cpu_float": __import__("torch").zeros(4)}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        w.save()                                                       # float pA must be digitised by the CUDA kernel
    assert get_seq_kit_and_flow_cell("dna-r9-prom") == ("SQK-LSK109", "FLO-PRO001")
    assert get_seq_kit_and_flow_cell("rna-004-prom") == ("sqk-rna004", "FLO-PRO004RA")
    with pytest.raises(ValueError):
        get_seq_kit_and_flow_cell("dna-r7-min")
    assert str(indexed_uuid(12)) == "00000000-0000-0000-0000-000000000012"


@pytest.mark.parametrize("ideal,preserve", [(True, False), (False, False), (False, True)])
def test_save_flat_equals_dict_save(tmp_path, ideal, preserve):
    """BLOW5Writer.save_flat (one contiguous int16 buffer + offsets, the read pipeline's fast path) writes byte-identical
    files to ``signals = {...}; save()``: numbering across batches, empty reads skipped, the per-record NumPy draws in
    the same order."""
    from seq2squiggle_b200.profiles import get_profile
    from seq2squiggle_b200.signal_io import BLOW5Writer
    prof = get_profile("dna-r10-prom")
    rng = np.random.default_rng(2)
    batches = []
    for b in range(3):
        lens = rng.integers(0, 400, size=17)
        lens[rng.integers(0, 17, size=3)] = 0                  # empty reads in the middle and possibly at the ends
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        flat = rng.integers(-300, 1200, size=int(off[-1])).astype(np.int16)
        batches.append(([f"read-{b}-{i}" for i in range(17)], flat, off))
    files = []
    for mode in ("dict", "flat"):
        w = BLOW5Writer(str(tmp_path / f"{mode}.blow5"), prof, ideal, "dna-r10-prom", preserve)
        np.random.seed(123)
        for names, flat, off in batches:
            if mode == "dict":
                w.signals = {n: flat[off[i]:off[i + 1]] for i, n in enumerate(names)}
                w.save()
            else:
                w.save_flat(names, flat, off)
        files.append(open(w.filename, "rb").read())
        assert w.reads_written == sum(int((np.diff(o) > 0).sum()) for _, _, o in batches)
    a, b = files
    # the header carries exp_start_time (wall clock, second resolution): compare everything after it
    ha = a.index(b"#char*"); hb = b.index(b"#char*")
    assert a[ha:] == b[hb:]
    recs = read_blow5(str(tmp_path / "flat.blow5"))["records"]
    assert [r["read_number"] for r in recs] == [i for i, l in enumerate(np.concatenate([np.diff(o) for _, _, o in batches])) if l > 0]


def test_svb_zd_known_answers_and_size(tmp_path):
    """svb-zd = zigzag-delta + StreamVByte (slow5lib's default signal compression), checked on hand-computed vectors:
    samples [3, 1, 1, -300, 32767, -32768] -> deltas [3, -2, 0, -301, 33067, -65535] -> zigzag [6, 3, 0, 601, 66134, 131069]
    -> byte lengths [1, 1, 1, 2 | 3, 3] -> control bytes 0b01000000, 0b00001010 -> 4 + 2 + 11 bytes; and a smooth signal
    (what a squiggle looks like) shrinks to 1.25 bytes per sample (one data byte + a quarter control byte)."""
    from seq2squiggle_b200 import _lib
    import ctypes as C
    prof = get_profile("dna-r10-prom")
    x = np.array([3, 1, 1, -300, 32767, -32768], dtype=np.int16)
    path = str(tmp_path / "k.blow5")
    w = BLOW5Writer(path, prof, True, "dna-r10-prom", True, record_compression="svb-zd")
    w.signals = {"kat": x}
    w.save()
    data = open(path, "rb").read()
    blob = bytes([6, 0, 0, 0, 0b01000000, 0b00001010, 6, 3, 0, 0x59, 0x02, 0x56, 0x02, 0x01, 0xFD, 0xFF, 0x01])
    assert (17).to_bytes(8, "little") + blob in data
    assert read_blow5(path)["records"][0]["signal"] == x.tolist()
    rng = np.random.default_rng(5)
    smooth = np.clip(np.cumsum(rng.integers(-40, 41, size=200000)) // 4 + 600, -32768, 32767).astype(np.int16)
    sizes = {}
    for comp in ("none", "svb-zd", "zlib+svb-zd"):
        p = str(tmp_path / f"s_{comp}.blow5")
        w = BLOW5Writer(p, prof, True, "dna-r10-prom", True, record_compression=comp)
        w.signals = {"s": smooth}
        w.save()
        sizes[comp] = os.path.getsize(p)
        assert read_blow5(p)["records"][0]["signal"] == smooth.tolist()
    assert sizes["svb-zd"] < 0.65 * sizes["none"] and sizes["zlib+svb-zd"] <= sizes["svb-zd"]
