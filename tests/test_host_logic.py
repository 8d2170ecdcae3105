"""CPU tests of the host side: the C-ABI libraries load and export every symbol the headers declare (no compute
call is made without a GPU), option plumbing (profiles, run options, CLI defaults), read batching / sharding and
the BLOW5 part merge of the multi-GPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s2s_[a-z0-9_]+)\s*\(", src)))


def test_cabi_exports_every_declared_symbol():
    from seq2squiggle_b200 import _lib
    _lib.build()
    _lib.build_blow5()
    for header, path, listed in (("s2s_b200.h", _lib.LIB_PATH, _lib.EXPORTS), ("s2s_blow5.h", _lib.BLOW5_LIB_PATH, _lib.EXPORTS_BLOW5)):
        declared = _declared(header)
        assert declared and sorted(listed) == declared, (header, set(declared) ^ set(listed))
        lib = ctypes.CDLL(path)
        for name in declared:
            assert hasattr(lib, name), f"{path} does not export {name}"
    lib = _lib.load()
    assert lib.s2s_abi_version() == 1
    assert lib.s2s_chunks_of_read(1000, 9) == 62 and lib.s2s_chunks_of_read(8, 9) == 0 and lib.s2s_chunks_of_read(9, 9) == 1
    cfg = _lib.S2SConfig(9, 2, 2, 1, 64, 256, 8, 16, 250, 165.0)
    assert lib.s2s_weights_count(ctypes.byref(cfg)) == 233_025 - 0 or lib.s2s_weights_count(ctypes.byref(cfg)) > 200_000
    bad = _lib.S2SConfig(9, 2, 2, 1, 128, 256, 8, 16, 250, 165.0)
    assert lib.s2s_weights_count(ctypes.byref(bad)) == -1 and b"unsupported architecture" in lib.s2s_last_error()


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    from seq2squiggle_b200.checkpoint import random_init_checkpoint, set_config
    from seq2squiggle_b200.engine import Engine
    cfg = set_config(None)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        Engine(random_init_checkpoint(cfg, 1)["state_dict"], cfg)


def test_weight_blob_matches_parameter_count():
    from seq2squiggle_b200 import _lib
    from seq2squiggle_b200.checkpoint import pack_weights, random_init_checkpoint, set_config
    for k in (9, 6):
        cfg = set_config(None)
        cfg["seq_kmer"] = k
        sd = random_init_checkpoint(cfg, 1)["state_dict"]
        blob = pack_weights(sd, cfg)
        c = _lib.S2SConfig(k, 2, 2, 1, 64, 256, 8, 16, 250, 165.0)
        assert blob.size == _lib.load().s2s_weights_count(ctypes.byref(c)) == sum(v.numel() for v in sd.values())
        with pytest.raises(KeyError):
            pack_weights({kk: v for kk, v in sd.items() if "w_qs" not in kk}, cfg)


def test_run_options_branch_mapping():
    """modules.py:410-432 / model.py:224-238 branch selection and inference.py:358-364 derived values."""
    from seq2squiggle_b200.engine import DUR_CONSTANT, DUR_NORMAL, DUR_SAMPLER, NOISE_OFF, NOISE_SAMPLER, NOISE_STATIC, RunOptions
    from seq2squiggle_b200.profiles import get_profile, update_config, update_profile
    for name, dwell in (("dna-r10-prom", 12.5), ("dna-r9-min", 4000 / 450), ("rna-004-min", 4000 / 130)):
        o = RunOptions.from_profile(get_profile(name), name, duration_sampling=False, dwell_std=0.0, noise_std=0.0)
        c = o.to_c(5)
        assert abs(c.dwell_mean - dwell) < 1e-5 and c.duration_mode == DUR_CONSTANT and c.noise_mode == NOISE_OFF
        assert c.rna_reverse == int(name.startswith("rna")) and c.chunk_id_base == 5
    p = get_profile("dna-r10-prom")
    assert RunOptions.from_profile(p, "dna-r10-prom", duration_sampling=False, dwell_std=1.5).to_c().duration_mode == DUR_NORMAL
    assert RunOptions.from_profile(p, "dna-r10-prom", duration_sampling=True, dwell_std=1.5).to_c().duration_mode == DUR_SAMPLER
    assert RunOptions.from_profile(p, "dna-r10-prom", noise_std=2.0, noise_sampling=False).to_c().noise_mode == NOISE_STATIC
    assert RunOptions.from_profile(p, "dna-r10-prom", noise_std=2.0, noise_sampling=True).to_c().noise_mode == NOISE_SAMPLER
    assert RunOptions.from_profile(p, "dna-r10-prom", noise_std=-1, noise_sampling=True).to_c().noise_mode == NOISE_OFF
    up = update_profile(get_profile("dna-r10-prom"), digitisation=4096, bps=None, range=300.0)
    assert up["digitisation"] == 4096 and up["bps"] == 400 and up["range"] == 300.0
    assert update_config("dna-r9-prom", {})["seq_kmer"] == 6 and update_config("rna-004-min", {})["seq_kmer"] == 9
    with pytest.raises(ValueError):
        update_config("dna-r7", {})


def test_cli_surface_matches_reference_defaults():
    from click.testing import CliRunner
    from seq2squiggle_b200.cli import main, predict
    opts = {o.name: o for o in predict.params}
    expect = dict(read_input=False, num_reads=-1, read_length=1000, coverage=-1, profile="dna-r10-prom",
                  noise_sampler=True, duration_sampler=True, dwell_mean=None, dwell_std=0.0, noise_std=2.0,
                  distr="expon", predict_batch_size=1024, export_every_n_samples=1000000, sample_rate=None, bps=None,
                  digitisation=None, range_val=None, offset_mean=None, offset_std=None, median_before_mean=None,
                  median_before_std=None, min_noise=0.0, min_duration=3, min_read_len=30, preserve_read_ids=False,
                  seed=0, model=None, config=None, verbosity="info")
    for name, default in expect.items():
        assert name in opts, name
        assert opts[name].default == default, (name, opts[name].default)
    assert "--noise-sampling" in opts["noise_sampler"].opts and "--noise-sampler" in opts["noise_sampler"].opts
    assert "--duration-sampling" in opts["duration_sampler"].opts
    assert opts["profile"].type.convert("dna_r9_min", None, None) == "dna-r9-min"     # README spelling
    r = CliRunner().invoke(main, ["predict"])
    assert r.exit_code == 1                                                          # seq2squiggle.py:513-515
    r = CliRunner().invoke(main, ["predict", "--show-advanced-options"])
    assert r.exit_code == 0 and "--min_duration" in r.output and "--range_val" in r.output
    assert CliRunner().invoke(main, ["train"]).exit_code != 0


def test_inference_run_boundary_errors(tmp_path):
    """Errors raised before any device work keep the reference's types/messages (inference.py:80-82, utils.py:257-262)."""
    from seq2squiggle_b200.checkpoint import check_model, set_config
    from seq2squiggle_b200.inference import get_saved_weights, get_writer
    from seq2squiggle_b200.profiles import get_profile
    p = get_profile("dna-r10-prom")
    with pytest.raises(ValueError, match=r"\.pod5, \.slow5, or \.blow5"):
        get_writer(str(tmp_path / "x.fast5"), p, True, 1000, "dna-r10-prom", False)
    existing = tmp_path / "sub" / "o.blow5"
    os.makedirs(existing.parent)
    existing.write_text("old")
    w, n = get_writer(str(existing), p, True, 1000, "dna-r10-prom", False)
    assert not existing.exists() and n == 1000 and type(w).__name__ == "BLOW5Writer"
    w, n = get_writer(str(tmp_path / "o.pod5"), p, True, 1000, "dna-r10-prom", False)
    assert n == float("inf") and type(w).__name__ == "POD5Writer"
    with pytest.raises(PermissionError):
        get_saved_weights("dna-r10-prom")
    cfg = set_config(None)
    other = dict(cfg, seq_kmer=6)
    with pytest.raises(ValueError, match="seq_kmer"):
        check_model(other, cfg)
    check_model(dict(cfg, dff=128), cfg)   # other mismatches only warn


def test_batching_and_sharding():
    from seq2squiggle_b200.inference import batch_reads, chunks_of_read, shard_reads
    rng = np.random.default_rng(0)
    lens = rng.integers(1, 4000, size=500)
    reads = [("A" * int(l), f"r{i}") for i, l in enumerate(lens)]
    batches = list(batch_reads(reads, 9, batch_chunks=2000))
    assert [x for b in batches for x in b] == reads                     # order kept, nothing split or lost
    sizes = [sum(chunks_of_read(len(s), 9) for s, _ in b) for b in batches]
    assert all(s >= 2000 for s in sizes[:-1]) and max(sizes) < 2000 + 250
    counts = [chunks_of_read(int(l), 9) for l in lens]
    for world in (1, 2, 3, 8):
        sh = shard_reads(counts, world)
        assert sh[0][0] == 0 and sh[-1][1] == len(counts)
        assert all(sh[i][1] == sh[i + 1][0] for i in range(world - 1))  # contiguous, disjoint, complete
        per = [sum(counts[a:b]) for a, b in sh]
        assert max(per) - min(per) <= 2 * max(counts)                 # each boundary is within one read of ideal
    assert shard_reads([], 4) == [(0, 0)] * 4
    assert shard_reads([5], 2) in ([(0, 0), (0, 1)], [(0, 1), (1, 1)])


def test_pack_reads_variants_agree():
    """Engine.pack_reads (torch tensors) and Engine.pack_reads_np (bytes + numpy, used by the staging pipeline) build the
    same byte stream and offsets for str, bytes and mixed inputs, including reads shorter than k and empty input."""
    from seq2squiggle_b200.engine import Engine
    reads = ["ACGTACGTACGTACGTACGTACGTA", "ACG", "", "N" * 40, "acgtacgtacgt" * 9]
    for inp in (reads, [r.encode() for r in reads], [reads[0], reads[1].encode(), reads[2], reads[3].encode(), reads[4]]):
        bases, ro, co = Engine.pack_reads(inp, 9)
        joined, ro2, co2 = Engine.pack_reads_np(inp, 9)
        assert bytes(bases.numpy().tobytes())[:len(joined)] == joined
        assert ro.tolist() == ro2.tolist() == [0, 25, 28, 28, 68, 176]
        assert co.tolist() == co2.tolist() == [0, 2, 2, 2, 4, 11]
    joined, ro2, co2 = Engine.pack_reads_np([], 9)
    assert joined == b"" and ro2.tolist() == [0] and co2.tolist() == [0]


def test_plan_batches():
    """Batches of a sharded run: contiguous, complete, a read is never split, every batch but the last reaches the
    target, and the rule equals predict_reads' own piece rule (one batch = one pipeline piece)."""
    from seq2squiggle_b200.inference import chunks_of_read, plan_batches
    rng = np.random.default_rng(1)
    counts = [chunks_of_read(int(l), 9) for l in rng.integers(1, 4000, size=700)]
    plan = plan_batches(counts, 2000)
    assert plan[0][0] == 0 and plan[-1][1] == len(counts) and all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
    sizes = [sum(counts[lo:hi]) for lo, hi in plan]
    assert all(s >= 2000 for s in sizes[:-1]) and max(sizes) < 2000 + 250
    # the piece rule of model.predict_reads: close the piece with the read that reaches the target
    ref, cur, n, lo = [], 0, 0, 0
    for i, c in enumerate(counts):
        n += c
        if n >= 2000:
            ref.append((lo, i + 1)); lo, n = i + 1, 0
    if lo < len(counts):
        ref.append((lo, len(counts)))
    assert plan == ref
    assert plan_batches([], 10) == [] and plan_batches([0, 0, 0], 10) == [(0, 3)] and plan_batches([50], 10) == [(0, 1)]


@pytest.mark.parametrize("samplers", [False, True])
def test_ordered_shared_file_equals_single_writer(tmp_path, samplers):
    """signal_io.SharedOrder + BLOW5Writer.begin_shared: three "ranks" (forked processes here, torchrun ranks in a run) write
    their round-robin batches into ONE file in whatever order their turn comes — a batch without reads, a batch whose
    reads are all empty, skipped (empty) reads inside a batch; with the samplers on every rank replays the single
    per-record NumPy draw stream.  Record bytes (and the end marker) equal one writer's."""
    from seq2squiggle_b200.profiles import get_profile
    from seq2squiggle_b200.signal_io import BLOW5Writer, SharedOrder
    from tests.blow5_reader import read_blow5, record_span
    rng = np.random.default_rng(4)
    prof = get_profile("dna-r10-prom")
    sigs = [rng.integers(-100, 900, size=0 if i in (2, 7, 8) else int(rng.integers(1, 300))).astype(np.int16)
            for i in range(23)]
    names = [f"r{i}" for i in range(23)]
    flat = np.concatenate(sigs)
    off = np.concatenate([[0], np.cumsum([len(x) for x in sigs])]).astype(np.int64)
    np.random.seed(11)
    single = BLOW5Writer(str(tmp_path / "single.blow5"), prof, not samplers, "dna-r10-prom", False)
    single.save_flat(names, flat, off)
    plan = [(0, 4), (4, 7), (7, 9), (9, 13), (13, 13), (13, 20), (20, 23)]     # batch 2: only empty reads; batch 4: no reads
    out = str(tmp_path / "o.blow5")
    world = 3
    SharedOrder(out + ".order", len(plan), create=True).close()

    def rank_main(r):              # one process per "rank" (its own NumPy stream, like a torchrun rank)
        np.random.seed(11)
        w = BLOW5Writer(out, prof, not samplers, "dna-r10-prom", False)
        w.begin_shared(SharedOrder(out + ".order", len(plan), create=False), r)
        for b in range(r, len(plan), world):
            lo, hi = plan[b]
            w.save_flat(names[lo:hi], flat, off[lo:hi + 1], tag=(b, lo))
        w.end_shared()

    import multiprocessing as mp
    ctx = mp.get_context("fork")
    procs = [ctx.Process(target=rank_main, args=(r,)) for r in (2, 1, 0)]           # rank 0 (the header) starts last
    for p_ in procs:
        p_.start()
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    a, b = read_blow5(str(tmp_path / "single.blow5")), read_blow5(out)
    assert a["records"] == b["records"] and len(a["records"]) == 20
    blobs = []
    for path in (str(tmp_path / "single.blow5"), out):
        lo, hi = record_span(path)
        blobs.append(open(path, "rb").read()[lo:])
    assert blobs[0] == blobs[1]


def test_default_config_equals_reference_yaml():
    """The packaged defaults (checkpoint.DEFAULT_CONFIG) carry the reference's config.yaml key for key, value for value,
    type for type (checked against the reference tree when it is mounted, against the golden checkpoint's stored config
    otherwise: oracle/make_golden.py wrote it from the reference YAML)."""
    import yaml
    import torch
    from seq2squiggle_b200.checkpoint import DEFAULT_CONFIG, set_config
    assert set_config(None) == DEFAULT_CONFIG and set_config(None) is not DEFAULT_CONFIG
    ref_yaml = "/root/reference/src/seq2squiggle/config.yaml"
    if os.path.exists(ref_yaml):
        ref = yaml.safe_load(open(ref_yaml))
    else:
        ck = torch.load(os.path.join(ROOT, "tests", "golden", "ckpt_k9_seed1.ckpt"), map_location="cpu", weights_only=False)
        ref = ck["hyper_parameters"]["config"]
    assert DEFAULT_CONFIG == ref
    assert all(type(DEFAULT_CONFIG[k]) is type(ref[k]) for k in ref)


def test_shipped_library_is_blackwell_native():
    """Static check of the built library (cuobjdump -sass, no GPU needed): the decoder attention and the fc + FFN kernels
    issue tcgen05.mma (UTCHMMA), read / write TMEM (LDTM / STTM), load by TMA (UTMALDG) and synchronise on mbarriers
    (SYNCS); the FFN block output leaves by a TMA store (UTMASTG); no legacy mma.sync (HMMA) exists anywhere."""
    import re
    import shutil
    import subprocess
    from collections import Counter
    from seq2squiggle_b200 import _lib
    if shutil.which("cuobjdump") is None or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not available")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    per_kernel, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per_kernel.setdefault(m.group(1), Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    assert sum(c["HMMA"] for c in per_kernel.values()) == 0

    def kernels(tag):
        found = [c for name, c in per_kernel.items() if tag in name]
        assert found, f"no kernel named *{tag}* in the library"
        return found

    for tag in ("k_tc_attn3", "k_tc_fc_ffn4", "k_tc_enc_attn"):
        for c in kernels(tag):
            assert c["UTCHMMA"] > 0 and c["LDTM"] > 0 and c["UTMALDG"] > 0 and c["SYNCS"] > 0, (tag, dict(c))
    assert all(c["STTM"] > 0 for c in kernels("k_tc_attn3") + kernels("k_tc_fc_ffn4"))   # fp16 P / hidden written back to TMEM
    assert any(c["UTMASTG"] > 0 for c in kernels("k_tc_fc_ffn4"))
    assert all(c["MUFU"] > 0 for c in kernels("k_tc_attn3"))
