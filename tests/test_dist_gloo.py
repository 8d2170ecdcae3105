"""world_size-2/3 CPU (gloo) tests of the multi-GPU host logic: the REAL ``inference_run`` (profile plumbing, writer
factory, ``get_reads_batches`` with its lengths-only sampler replay, ``plan_batches``, round-robin batches, ordered
writes of all ranks into ONE shared BLOW5 file through ``signal_io.SharedOrder``) with only the model replaced by a
deterministic stand-in keyed, like the device path, by the GLOBAL chunk index it is handed.  The file of an N-rank run
must equal the single-process file record for record and byte for byte (records + end marker; the header carries the
wall-clock ``exp_start_time`` of each run) — including, with the samplers on, the per-record ``offset`` /
``median_before`` NumPy draws, which every rank replays from the single-process stream."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


RUN_WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    sys.path.insert(0, %(root)r)
    from seq2squiggle_b200 import inference, model as model_mod
    from seq2squiggle_b200.checkpoint import DEFAULT_CONFIG
    from seq2squiggle_b200.cli import set_seeds

    model_mod.PIPE_CHUNKS = int(os.environ.get("S2S_TEST_PIPE", 300))   # small batches: about a dozen per run, round-robin

    class FakeModel:
        '''What inference_run touches of seq2squiggle: load_from_checkpoint, hparams.config, chunks_done, predict_reads,
        on_predict_epoch_end.  A read's signal depends on (global index of its first chunk, its length) only.'''
        def __init__(self, writer):
            self.out_writer, self.chunks_done = writer, 0
            self.hparams = model_mod._HParams(config=dict(DEFAULT_CONFIG))

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, out_writer=None, **kw):
            return cls(out_writer)

        def predict_reads(self, reads, chunk_id_base=None, tag=None):
            if chunk_id_base is not None:
                self.chunks_done = chunk_id_base
            names, sigs = [], []
            for seq, name in reads:
                n = max(len(seq) - 9 + 1, 0)
                n = -(-n // 16)
                g = np.random.default_rng([self.chunks_done, len(seq)])
                k = 0 if len(seq) %% 7 == 0 else n * 40           # some reads come out empty (skipped records)
                sigs.append(g.integers(-500, 1500, size=k).astype(np.int16))
                names.append(name)
                self.chunks_done += n
            off = np.concatenate([[0], np.cumsum([len(s) for s in sigs])]).astype(np.int64)
            flat = np.concatenate(sigs) if sigs else np.zeros(0, np.int16)
            if tag is not None:
                self.out_writer.save_flat(names, flat, off, tag=tag)
            else:
                self.out_writer.save_flat(names, flat, off)

        def on_predict_epoch_end(self):
            pass

    model_mod.seq2squiggle = FakeModel
    torch.cuda.set_device = lambda *_a, **_k: None
    fasta, out, mode, samplers, seed = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4] == "1", int(sys.argv[5])
    seed = set_seeds(seed)
    inference.inference_run(config=dict(DEFAULT_CONFIG), saved_weights="unused.ckpt", fasta=fasta, read_input=(mode == "read"),
                            n=90, r=700, c=-1, out=out, profile="dna-r10-prom", dwell_mean=None, dwell_std=0.0,
                            noise_std=0.0, noise_sampling=False, duration_sampling=samplers, distr="expon",
                            predict_batch_size=1024, export_every_n_samples=2000000, sample_rate=None, bps=None,
                            digitisation=None, range_val=None, offset_mean=None, offset_std=None, median_before_mean=None,
                            median_before_std=None, min_noise=0.0, min_duration=3, min_read_len=30,
                            preserve_read_ids=False, seed=seed)
    print("SEED", seed)
""")


def _run(tmp_path, script, fasta, out, mode, samplers, seed, world, comp="none"):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["S2S_BLOW5_COMPRESS"] = comp
    args = [sys.executable, str(script), str(fasta), str(out), mode, "1" if samplers else "0", str(seed)]
    if world == 1:
        return [subprocess.run(args, check=True, env=env, timeout=300, capture_output=True, text=True).stdout]
    port = _free_port()
    procs = []
    for r in range(world):
        e = dict(env, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen(args, env=e, stdout=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        o, _ = p.communicate(timeout=300)
        assert p.returncode == 0
        outs.append(o)
    return outs


def _inputs(tmp_path, mode):
    import numpy as np
    rng = np.random.default_rng(2)
    fasta = tmp_path / "in.fasta"
    if mode == "reference":
        contigs = ["".join(rng.choice(list("ACGTN"), n, p=[0.245, 0.245, 0.245, 0.245, 0.02])) for n in (15000, 9000)]
    else:
        contigs = ["".join(rng.choice(list("ACGT"), int(n))) for n in rng.integers(200, 3000, size=12)]
    fasta.write_text("".join(f">s{i}\n{g}\n" for i, g in enumerate(contigs)))
    script = tmp_path / "run_worker.py"
    script.write_text(RUN_WORKER % {"root": ROOT})
    return fasta, script


@pytest.mark.parametrize("mode,world,samplers,comp", [("reference", 2, False, "none"), ("reference", 3, True, "zlib+svb-zd"),
                                                      ("read", 2, True, "none")])
def test_inference_run_n_ranks_equal_one_rank_on_cpu(tmp_path, mode, world, samplers, comp):
    """``comp``: pyslow5's default pair (zlib records + svb-zd signal) works in a sharded run too — every rank knows its
    records' ``start_time`` before it encodes them, so nothing inside a compressed record has to be patched later."""
    from tests.blow5_reader import read_blow5, record_span
    fasta, script = _inputs(tmp_path, mode)
    _run(tmp_path, script, fasta, tmp_path / "one.blow5", mode, samplers, 21, 1, comp)
    _run(tmp_path, script, fasta, tmp_path / "many.blow5", mode, samplers, 21, world, comp)
    a, b = read_blow5(str(tmp_path / "one.blow5")), read_blow5(str(tmp_path / "many.blow5"))
    assert a["record_compression"] == b["record_compression"] == (1 if "zlib" in comp else 0)
    assert a["signal_compression"] == b["signal_compression"] == (1 if "svb-zd" in comp else 0)
    assert 40 < len(a["records"]) < 90                 # some reads produced no signal and were skipped
    assert a["records"] == b["records"]                # incl. read_number, start_time, offset, median_before
    if samplers:
        assert len({r["offset"] for r in a["records"]}) > 10     # per-record draws, not the profile constant
    blobs = []
    for name in ("one.blow5", "many.blow5"):
        lo, hi = record_span(str(tmp_path / name))
        blobs.append(open(tmp_path / name, "rb").read()[lo:])
    assert blobs[0] == blobs[1]
    # nothing but the output is left behind: no part files, no hand-over table
    assert sorted(os.listdir(tmp_path)) == ["in.fasta", "many.blow5", "one.blow5", "run_worker.py"]


def test_more_ranks_than_batches(tmp_path, monkeypatch):
    """A run that is one batch long under three ranks: ranks 1 and 2 own nothing, open the shared file, write nothing and
    leave; the file is the single-process file."""
    from tests.blow5_reader import read_blow5, record_span
    monkeypatch.setenv("S2S_TEST_PIPE", "1000000")
    fasta, script = _inputs(tmp_path, "reference")
    _run(tmp_path, script, fasta, tmp_path / "one.blow5", "reference", True, 21, 1)
    _run(tmp_path, script, fasta, tmp_path / "many.blow5", "reference", True, 21, 3)
    a, b = read_blow5(str(tmp_path / "one.blow5")), read_blow5(str(tmp_path / "many.blow5"))
    assert a["records"] == b["records"] and len(a["records"]) > 40
    assert sorted(os.listdir(tmp_path)) == ["in.fasta", "many.blow5", "one.blow5", "run_worker.py"]


def test_random_seed_is_shared_by_all_ranks(tmp_path):
    """``-s 0`` (the CLI default) draws a random seed: under torchrun rank 0 draws it and broadcasts it, so that every
    rank derives the same read list / batches / Philox key (cli.resolve_random_seed).  Two 2-rank runs: each run's ranks
    agree, the file is complete and consistent, and the two runs (different seeds) differ."""
    from tests.blow5_reader import read_blow5
    fasta, script = _inputs(tmp_path, "reference")
    seeds = []
    for name in ("a.blow5", "b.blow5"):
        outs = _run(tmp_path, script, fasta, tmp_path / name, "reference", False, 0, 2)
        got = {line.split()[1] for o in outs for line in o.splitlines() if line.startswith("SEED")}
        assert len(got) == 1, got
        seeds.append(got.pop())
        recs = read_blow5(str(tmp_path / name))["records"]
        nums = [r["read_number"] for r in recs]
        assert nums == sorted(set(nums)) and len(recs) > 40          # no duplicate / missing read numbers
        t = 0
        for r in recs:
            assert r["start_time"] == t
            t += r["len_raw_signal"]
    assert seeds[0] != seeds[1]
