"""GPU tests of the reference-facing plug points (model.seq2squiggle, inference_run, the CLI) — everything goes
through libs2s_b200.so; the oracle only checks."""
import os
import random
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import s2s_oracle as orc
from oracle.profiles_kat import PROFILES
from tests.blow5_reader import read_blow5

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ckpt(golden_dir, name="ckpt_k9_seed1.ckpt"):
    path = os.path.join(golden_dir, name)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    return path, ck["state_dict"], ck["hyper_parameters"]["config"]


def _writer(tmp_path, profile="dna-r10-prom", ideal=True, name="o.blow5", preserve=True):
    from seq2squiggle_b200.profiles import get_profile
    from seq2squiggle_b200.signal_io import BLOW5Writer
    return BLOW5Writer(str(tmp_path / name), get_profile(profile), ideal, profile, preserve)


def _oracle_signals(sd, cfg, reads, profile, **opts):
    prof = PROFILES[profile]
    out = {}
    for seq, name in reads:
        data = orc.split_sequence(seq, cfg)
        if not data.size:
            continue
        pa = orc.predict_step(sd, cfg, torch.from_numpy(data), **opts)
        sig = orc.assemble_reads([name] * len(data), pa)[name].reshape(-1).numpy()
        out[name] = (pa.numpy(), orc.digitise(sig, prof["digitisation"], prof["range"], prof["offset_mean"],
                                              rna=profile.startswith("rna")))
    return out


def _reads(n, seed=0, lo=20, hi=700):
    rng = np.random.default_rng(seed)
    return [("".join(rng.choice(list("ACGT"), int(rng.integers(lo, hi)))), f"read_{i}") for i in range(n)]


def test_predict_step_and_export_match_oracle(golden_dir, tmp_path):
    """model.py:195-307 through the plug point: one-hot DataLoader batches of 64 chunks that cut reads in the middle,
    periodic export with keep_last, final export; fp32 parity path so the zero-strip lengths agree."""
    from seq2squiggle_b200.model import seq2squiggle
    path, sd, cfg = _ckpt(golden_dir)
    w = _writer(tmp_path)
    m = seq2squiggle.load_from_checkpoint(path, out_writer=w, dwell_mean=12.5, dwell_std=0.0, noise_std=0.0,
                                          noise_sampling=False, duration_sampling=False, export_every_n_samples=150,
                                          min_noise=0.0, min_duration=3, precision="fp32")
    assert m.hparams.config["seq_kmer"] == 9 and m.hparams["dwell_mean"] == 12.5
    reads = _reads(14, seed=1)
    ids, chunks = [], []
    for seq, name in reads:
        c = orc.split_sequence(seq, cfg)
        if c.size:
            chunks.append(c)
            ids += [name] * len(c)
    data = torch.from_numpy(np.concatenate(chunks, 0))
    for b in range(0, len(ids), 64):
        m.predict_step((ids[b:b + 64], data[b:b + 64]))
    m.on_predict_epoch_end()
    ref = _oracle_signals(sd, cfg, reads, "dna-r10-prom", dwell_mean=12.5, min_duration=3)
    f = read_blow5(w.filename)
    got = {r["read_id"]: np.array(r["signal"], dtype=np.int16) for r in f["records"]}
    assert list(got) == [n for _, n in reads if n in ref]                 # export order = first-seen order
    exact = 0
    for name, (pa, raw) in ref.items():
        if len(got[name]) == len(raw):
            exact += 1
            assert np.abs(got[name].astype(int) - raw.astype(int)).max() <= 1
    assert exact >= len(ref) - 1                                          # a ReLU sign flip may change one length
    assert [r["start_time"] for r in f["records"]] == list(np.cumsum([0] + [len(got[n]) for n in got])[:-1])


def test_predict_reads_equals_predict_step(golden_dir, tmp_path):
    """The native fast path (bytes in, int16 out) and the DataLoader-batch path give identical signals, samplers on."""
    from seq2squiggle_b200.model import seq2squiggle
    path, sd, cfg = _ckpt(golden_dir)
    reads = _reads(20, seed=2)
    kw = dict(dwell_mean=12.5, dwell_std=0.0, noise_std=2.0, noise_sampling=True, duration_sampling=True,
              export_every_n_samples=10 ** 9, min_noise=0.0, min_duration=3, seed=11)
    w1 = _writer(tmp_path, ideal=False, name="a.blow5")
    m1 = seq2squiggle.load_from_checkpoint(path, out_writer=w1, **kw)
    m1.predict_reads(reads[:9])
    m1.predict_reads(reads[9:])
    m1.on_predict_epoch_end()
    w2 = _writer(tmp_path, ideal=False, name="b.blow5")
    m2 = seq2squiggle.load_from_checkpoint(path, out_writer=w2, **kw)
    ids, chunks = [], []
    for seq, name in reads:
        c = orc.split_sequence(seq, cfg)
        if c.size:
            chunks.append(c)
            ids += [name] * len(c)
    data = torch.from_numpy(np.concatenate(chunks, 0))
    for b in range(0, len(ids), 50):
        m2.predict_step((ids[b:b + 50], data[b:b + 50]))
    m2.on_predict_epoch_end()
    a, b = read_blow5(w1.filename), read_blow5(w2.filename)
    assert [r["read_id"] for r in a["records"]] == [r["read_id"] for r in b["records"]]
    for ra, rb in zip(a["records"], b["records"]):
        assert ra["signal"] == rb["signal"]
    assert sum(len(r["signal"]) for r in a["records"]) > 1000


def test_read_pipeline_edge_cases_and_errors(golden_dir, tmp_path):
    """The staging pipeline behind predict_reads: empty batches, batches with no chunk at all, growth of the recycled
    slots when a later batch is much larger, reuse across epochs, and a writer exception surfacing at the next call."""
    from seq2squiggle_b200.model import seq2squiggle
    path, sd, cfg = _ckpt(golden_dir)
    kw = dict(dwell_mean=12.5, dwell_std=0.0, noise_std=2.0, noise_sampling=True, duration_sampling=True,
              export_every_n_samples=10 ** 9, min_noise=0.0, min_duration=3, seed=5)
    w = _writer(tmp_path, ideal=True, name="edge.blow5")
    m = seq2squiggle.load_from_checkpoint(path, out_writer=w, **kw)
    small = [(s, "small_" + n) for s, n in _reads(6, seed=3, lo=40, hi=200)]
    big = [(s, "big_" + n) for s, n in _reads(40, seed=4, lo=2000, hi=6000)]
    m.predict_reads([])
    m.predict_reads([("ACG", "too-short"), ("", "empty")])
    m.predict_reads(small)
    m.predict_reads(big)                      # far larger than the first batch: every slot has to grow
    m.on_predict_epoch_end()
    m.predict_reads(small[:2])                # second epoch on the same pipeline
    m.on_predict_epoch_end()
    f = read_blow5(w.filename)
    assert [r["read_id"] for r in f["records"]] == [n for _, n in small + big + small[:2]]
    assert [r["read_number"] for r in f["records"]] == list(range(2, 2 + 48))          # the two chunk-less reads count

    class Boom:
        profile, profile_name, is_rna = w.profile, "dna-r10-prom", False

        def save(self):
            raise OSError("disk full")
    m2 = seq2squiggle.load_from_checkpoint(path, out_writer=Boom(), **kw)
    m2.predict_reads(small)
    with pytest.raises(OSError, match="disk full"):
        m2.on_predict_epoch_end()
    m2.out_writer = _writer(tmp_path, ideal=True, name="after.blow5")
    m2.predict_reads(small)                   # the pipeline is usable again after the error was reported
    m2.on_predict_epoch_end()
    assert len(read_blow5(m2.out_writer.filename)["records"]) == len(small)


def test_writer_digitises_float_pa_on_device(golden_dir, tmp_path):
    """Reference writer contract: signals = {read_id: float pA tensor}; digitised by the CUDA kernel (signal_io.py:134-141)."""
    fx = np.load(os.path.join(golden_dir, "digitise_kat.npz"))
    for pname in ("dna-r10-prom", "rna-004-min"):
        w = _writer(tmp_path, profile=pname, name=f"{pname}.blow5")
        pa = fx[pname + "/pa"]
        w.signals = {"x": torch.from_numpy(pa).cuda(), "empty": torch.zeros(0).cuda()}
        w.save()
        rec = read_blow5(w.filename)["records"]
        assert len(rec) == 1 and rec[0]["signal"] == fx[pname + "/raw"].tolist()


@pytest.mark.parametrize("profile,ckpt", [("dna-r10-prom", "ckpt_k9_seed1.ckpt"), ("dna_r9_min", "ckpt_k6_seed2.ckpt")])
def test_cli_predict_reference_mode_deterministic(golden_dir, tmp_path, profile, ckpt):
    """BASELINE config 1 in miniature, through the command line: reference mode, -n 40, deterministic options."""
    from seq2squiggle_b200.reads import get_reads
    path, sd, cfg = _ckpt(golden_dir, ckpt)
    rng = np.random.default_rng(7)
    fasta = tmp_path / "genome.fasta"
    g = "".join(rng.choice(list("ACGTN"), 9000, p=[0.245, 0.245, 0.245, 0.245, 0.02]))
    fasta.write_text(">chr1 test\n" + "\n".join(g[i:i + 70] for i in range(0, len(g), 70)) + "\n")
    out = tmp_path / "sim.blow5"
    cmd = [sys.executable, "-m", "seq2squiggle_b200", "predict", str(fasta), "-o", str(out), "-m", path, "-n", "40",
           "-r", "400", "-s", "5", "--profile", profile, "--duration-sampling", "False", "--noise-sampling", "False",
           "--noise-std", "0", "--dwell-std", "0", "--precision", "fp32"]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "Prediction done." in res.stderr
    pname = profile.replace("_", "-")
    random.seed(5)
    reads = list(get_reads(str(fasta), False, 40, 400, -1, dict(cfg), "expon", 5, pname, 30)[0])
    reads = [(s, f"r{i}") for i, (s, _) in enumerate(reads)]
    dwell = PROFILES[pname]["sample_rate"] / PROFILES[pname]["bps"]
    ref = _oracle_signals(sd, cfg, reads, pname, dwell_mean=dwell, min_duration=3)
    f = read_blow5(str(out))
    assert f["attrs"]["sample_frequency"] == str(PROFILES[pname]["sample_rate"])
    assert len(f["records"]) == len(ref) > 30
    same = 0
    for rec, (name, (pa, raw)) in zip(f["records"], ref.items()):
        got = np.array(rec["signal"], dtype=np.int16)
        if len(got) == len(raw):
            same += 1
            assert np.abs(got.astype(int) - raw.astype(int)).max() <= 1
        assert rec["offset"] == PROFILES[pname]["offset_mean"]            # ideal mode: fixed record metadata
    assert same >= len(ref) - 2


def test_cli_two_ranks_equal_one_rank(golden_dir, tmp_path):
    """`torchrun --nproc-per-node 2 -m seq2squiggle_b200 predict ...` (both ranks on this GPU) writes the same records as
    the single-process command: the batches of the read list are dealt to the ranks round-robin, Philox is keyed by the
    global chunk index, and both ranks write their records into the one output file in read order (read ids,
    read_number, start_time AND the per-record offset / median_before draws equal the single-process file's)."""
    path, sd, cfg = _ckpt(golden_dir)
    rng = np.random.default_rng(17)
    fasta = tmp_path / "genome.fasta"
    g = "".join(rng.choice(list("ACGT"), 20000))
    fasta.write_text(">chr1\n" + "\n".join(g[i:i + 70] for i in range(0, len(g), 70)) + "\n")
    args = [str(fasta), "-m", path, "-n", "300", "-r", "500", "-s", "5", "--profile", "dna-r10-prom", "--noise-std", "2.0"]
    one, two = tmp_path / "one.blow5", tmp_path / "two.blow5"
    env = dict(os.environ, S2S_PIPE_CHUNKS="1500")      # about seven batches: both ranks get several
    res = subprocess.run([sys.executable, "-m", "seq2squiggle_b200", "predict", *args, "-o", str(one)], cwd=ROOT,
                         capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", "-m", "seq2squiggle_b200", "predict",
                          *args, "-o", str(two)], cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    assert "of 7 batches" in res.stderr or "batches" in res.stderr
    a, b = read_blow5(str(one)), read_blow5(str(two))
    assert len(a["records"]) == len(b["records"]) > 250
    assert sorted(p.name for p in tmp_path.iterdir()) == ["genome.fasta", "one.blow5", "two.blow5"]   # no parts, no table
    assert len({r["offset"] for r in a["records"]}) > 100           # samplers on: per-record draws
    for i, (x, y) in enumerate(zip(a["records"], b["records"])):
        assert x == y, i
        assert y["read_number"] == i


def test_inference_run_read_mode_samplers_on(golden_dir, tmp_path):
    """Read mode, default samplers, tensor-core path: runs, every read produces a plausible record, and a second run
    with the same seed reproduces the file's signals bit for bit."""
    from seq2squiggle_b200.checkpoint import set_config
    from seq2squiggle_b200.cli import set_seeds
    from seq2squiggle_b200.inference import inference_run
    path, sd, cfg = _ckpt(golden_dir)
    fasta = tmp_path / "reads.fasta"
    reads = _reads(30, seed=4, lo=5, hi=3000)                              # includes reads shorter than k
    fasta.write_text("".join(f">{n}\n{s}\n" for s, n in reads))
    outs = []
    for i in range(2):
        out = tmp_path / f"o{i}.blow5"
        set_seeds(21)
        inference_run(config=set_config(None), saved_weights=path, fasta=str(fasta), read_input=True, n=-1, r=1000, c=-1,
                      out=str(out), profile="dna-r10-prom", dwell_mean=None, dwell_std=0.0, noise_std=2.0,
                      noise_sampling=True, duration_sampling=True, distr="expon", predict_batch_size=1024,
                      export_every_n_samples=1000000, sample_rate=None, bps=None, digitisation=None, range_val=None,
                      offset_mean=None, offset_std=None, median_before_mean=None, median_before_std=None, min_noise=0.0,
                      min_duration=3, min_read_len=30, preserve_read_ids=True, seed=21)
        outs.append(read_blow5(str(out)))
    a, b = outs
    long_enough = [n for s, n in reads if len(s) >= 9]
    assert [r["read_id"] for r in a["records"]] == long_enough
    for ra, rb in zip(a["records"], b["records"]):
        assert ra["signal"] == rb["signal"] and ra["offset"] == rb["offset"] and ra["median_before"] == rb["median_before"]
    lens = {n: len(s) for s, n in reads}
    for r in a["records"]:
        nch = orc.n_chunks_of_read(lens[r["read_id"]], 9)
        assert 0 < r["len_raw_signal"] <= 250 * nch
