"""world_size-2 CPU (gloo) test of the multi-GPU host logic of inference_run: every rank derives the same read list
from the seed, takes its shard_reads() range, writes a BLOW5 part, rank 0 merges after the barrier.  The device call
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
    from seq2squiggle_b200.inference import chunks_of_read, get_writer, merge_blow5_parts, part_path, shard_reads
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
    w.signals = sig
    w.save()
    if world > 1:
        dist.barrier()
        if rank == 0:
            merge_blow5_parts(out, [part_path(out, i) for i in range(world)], False)
        dist.barrier()
        dist.destroy_process_group()
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_sharded_run_equals_single_process(tmp_path):
    from tests.blow5_reader import read_blow5
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    subprocess.run([sys.executable, str(script), str(tmp_path / "one.blow5")], check=True, env=env, timeout=300)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), str(tmp_path / "two.blow5")], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    a, b = read_blow5(str(tmp_path / "one.blow5")), read_blow5(str(tmp_path / "two.blow5"))
    assert len(a["records"]) == len(b["records"]) > 40
    assert a["records"] == b["records"]
