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
