"""world_size-2 CPU (gloo) test of the multi-GPU host logic of inference_run: every rank derives the same read list
from the seed, takes its shard_reads() range, writes a BLOW5 part with global read numbers, and all ranks splice
their parts into the output in parallel (splice_parts_collective).  The device call
is replaced by a deterministic stand-in keyed by the GLOBAL chunk index (exactly what the Philox keying guarantees
on the GPU), so the merged file must equal the single-process file record for record."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, random, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, %(root)r)
    from seq2squiggle_b200.inference import chunks_of_read, get_writer, part_path, shard_reads, splice_parts_collective
    from seq2squiggle_b200.profiles import get_profile
    from seq2squiggle_b200.reads import sampling
    from seq2squiggle_b200.signal_io import BLOW5Writer

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    out = sys.argv[1]
    if world > 1:
        dist.init_process_group("gloo")
    rng = np.random.default_rng(5)
    genome = "".join(rng.choice(list("ACGT"), 20000))
    random.seed(9)
    reads = sampling(60, [genome], [len(genome)], 600, 9, len(genome), "expon", "dna-r10-prom", 30)
    counts = [chunks_of_read(len(s), 9) for s in reads]
    lo, hi = shard_reads(counts, world)[rank]
    base = sum(counts[:lo])

    def fake_device(seq, first_chunk):            # stand-in for s2s_forward_reads: depends on global chunk ids only
        n = chunks_of_read(len(seq), 9)
        g = np.random.default_rng([first_chunk, n])
        return g.integers(-500, 1500, size=n * 100).astype(np.int16)

    sig, c = {}, base
    for i in range(lo, hi):
        sig[f"read{i}"] = fake_device(reads[i], c)
        c += counts[i]
    prof = get_profile("dna-r10-prom")
    path = out if world == 1 else part_path(out, rank)
    w, _ = get_writer(path, prof, True, 1000000, "dna-r10-prom", False)   # the same factory (and extension check) as inference_run
    w._id_base = lo                                # as inference_run: global read numbers / ids from the start
    w.signals = sig
    w.save()
    if world > 1:
        splice_parts_collective(out, path, rank, world, w.samples_written, dist)
        assert not os.path.exists(path)
        dist.barrier()
        dist.destroy_process_group()
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


import pytest


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_run_equals_single_process(tmp_path, world, monkeypatch):
    if world == 3:      # parts of the ranks > 0 in another directory (S2S_PART_DIR, e.g. a tmpfs)
        (tmp_path / "parts").mkdir()
        monkeypatch.setenv("S2S_PART_DIR", str(tmp_path / "parts"))
    from tests.blow5_reader import read_blow5
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    subprocess.run([sys.executable, str(script), str(tmp_path / "one.blow5")], check=True, env=env, timeout=300)
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), str(tmp_path / "many.blow5")], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    a, b = read_blow5(str(tmp_path / "one.blow5")), read_blow5(str(tmp_path / "many.blow5"))
    assert len(a["records"]) == len(b["records"]) > 40
    assert a["records"] == b["records"]
    # byte-identical records and one end marker (the header carries the wall-clock exp_start_time of each run)
    from seq2squiggle_b200.inference import blow5_record_span
    blobs = []
    for name in ("one.blow5", "many.blow5"):
        lo, hi = blow5_record_span(str(tmp_path / name))
        data = open(tmp_path / name, "rb").read()
        assert len(data) == hi + 5
        blobs.append(data[lo:])
    assert blobs[0] == blobs[1]
    assert sorted(os.listdir(tmp_path)) == sorted(["worker.py", "one.blow5", "many.blow5"] + (["parts"] if world == 3 else []))
    assert world != 3 or os.listdir(tmp_path / "parts") == []


# ---------------------------------------------------------------------------------------------------
# The REAL inference_run (profile plumbing, writer factory, get_reads / get_reads_shard, shard numbering, parallel splice)
# on CPU: only the model is a stand-in, keyed like the device path by the global chunk index it is handed.
# ---------------------------------------------------------------------------------------------------
RUN_WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    sys.path.insert(0, %(root)r)
    from seq2squiggle_b200 import inference, model as model_mod
    from seq2squiggle_b200.checkpoint import DEFAULT_CONFIG
    from seq2squiggle_b200.cli import set_seeds

    class FakeModel:
        '''What inference_run touches of seq2squiggle: load_from_checkpoint, hparams.config, chunks_done, predict_reads,
        on_predict_epoch_end.  A read's signal depends on (global index of its first chunk, its length) only.'''
        def __init__(self, writer):
            self.out_writer, self.chunks_done = writer, 0
            self.hparams = model_mod._HParams(config=dict(DEFAULT_CONFIG))

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, out_writer=None, **kw):
            return cls(out_writer)

        def predict_reads(self, reads):
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
            self.out_writer.save_flat(names, flat, off)

        def on_predict_epoch_end(self):
            pass

    model_mod.seq2squiggle = FakeModel
    torch.cuda.set_device = lambda *_a, **_k: None
    fasta, out, mode = sys.argv[1], sys.argv[2], sys.argv[3]
    set_seeds(21)
    inference.inference_run(config=dict(DEFAULT_CONFIG), saved_weights="unused.ckpt", fasta=fasta, read_input=(mode == "read"),
                            n=90, r=700, c=-1, out=out, profile="dna-r10-prom", dwell_mean=None, dwell_std=0.0,
                            noise_std=0.0, noise_sampling=False, duration_sampling=False, distr="expon",
                            predict_batch_size=1024, export_every_n_samples=2000000, sample_rate=None, bps=None,
                            digitisation=None, range_val=None, offset_mean=None, offset_std=None, median_before_mean=None,
                            median_before_std=None, min_noise=0.0, min_duration=3, min_read_len=30,
                            preserve_read_ids=False, seed=21)
""")


@pytest.mark.parametrize("mode", ["reference", "read"])
def test_inference_run_two_ranks_equal_one_rank_on_cpu(tmp_path, mode):
    import numpy as np
    from seq2squiggle_b200.inference import blow5_record_span
    from tests.blow5_reader import read_blow5
    rng = np.random.default_rng(2)
    fasta = tmp_path / "in.fasta"
    if mode == "reference":
        contigs = ["".join(rng.choice(list("ACGTN"), n, p=[0.245, 0.245, 0.245, 0.245, 0.02])) for n in (15000, 9000)]
    else:
        contigs = ["".join(rng.choice(list("ACGT"), int(n))) for n in rng.integers(200, 3000, size=12)]
    fasta.write_text("".join(f">s{i}\n{g}\n" for i, g in enumerate(contigs)))
    script = tmp_path / "run_worker.py"
    script.write_text(RUN_WORKER % {"root": ROOT})
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    subprocess.run([sys.executable, str(script), str(fasta), str(tmp_path / "one.blow5"), mode], check=True, env=env,
                   timeout=300)
    port = _free_port()
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), str(fasta), str(tmp_path / "two.blow5"), mode], env=e))
    for p in procs:
        assert p.wait(timeout=300) == 0
    a, b = read_blow5(str(tmp_path / "one.blow5")), read_blow5(str(tmp_path / "two.blow5"))
    assert 40 < len(a["records"]) < 90                 # some reads produced no signal and were skipped
    assert a["records"] == b["records"]
    blobs = []
    for name in ("one.blow5", "two.blow5"):
        lo, hi = blow5_record_span(str(tmp_path / name))
        blobs.append(open(tmp_path / name, "rb").read()[lo:])
    assert blobs[0] == blobs[1]
    assert sorted(os.listdir(tmp_path)) == ["in.fasta", "one.blow5", "run_worker.py", "two.blow5"]
