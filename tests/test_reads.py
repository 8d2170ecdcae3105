"""Host read source (seq2squiggle_b200/reads.py) against the reference's own sampler output
(tests/golden/read_sampling.json, produced by utils.sampling via oracle/make_golden.py) and against scipy."""
import gzip
import hashlib
import json
import os
import random

import numpy as np
import pytest

from seq2squiggle_b200 import reads as R


def test_sampling_matches_reference_golden(golden_dir):
    fx = json.load(open(os.path.join(golden_dir, "read_sampling.json")))
    genome = fx["genome"]
    assert len(fx["cases"]) >= 3
    for case in fx["cases"]:
        random.seed(case["seed"])
        got = R.sampling(case["n"], [genome], [len(genome)], case["r"], case["seed"], len(genome), case["distr"],
                         case["profile"], 30)
        assert [len(r) for r in got] == case["lens"], case["distr"]
        assert [hashlib.md5(r.encode()).hexdigest() for r in got] == case["md5"], case["distr"]


def test_length_draws_equal_scipy():
    st = pytest.importorskip("scipy.stats")
    for seed in (0, 1, 7, 12345, 2 ** 31 - 5):
        e = st.expon.rvs(loc=213.98910256668592, scale=6972.5319847131141, size=1, random_state=seed)
        assert R.draw_expon_dis(1000, seed, 10 ** 7) == np.clip((e[0] * 1000 / 7106.0).astype(int), 1, 10 ** 7)
        g = st.gamma.rvs(6.3693711, 0.53834893, size=1, random_state=seed)
        assert R.draw_gamma_dis(1000, seed, 10 ** 7) == np.clip(int((g * 1000 / 4.39)[0]), 1, 10 ** 7)
        b = st.beta.rvs(1.778, 7.892, 316.758, 34191.257, size=1, random_state=seed)
        assert R.draw_beta_dis(1000, seed, 10 ** 7) == np.clip((b[0] * 1000 / 6615.0).astype(int), 1, 10 ** 7)


def test_vectorised_first_draw_equals_randomstate():
    """The batched read-length path (one init_genrand recurrence for all seeds) is bit-identical to constructing one
    numpy RandomState per read, which is what scipy's rvs(random_state=int) does in the reference (utils.py:325-331)."""
    from seq2squiggle_b200 import reads as R
    seeds = np.concatenate([np.arange(0, 500), np.random.default_rng(1).integers(0, 2 ** 32 - 1, size=500),
                            [2 ** 32 - 1, 2 ** 31, 2 ** 31 - 1]]).astype(np.uint64)
    ref = np.array([np.random.RandomState(int(s)).random_sample() for s in seeds])
    assert np.array_equal(R.mt19937_first_doubles(seeds), ref)
    for mean, total in ((1000, 48502), (5000, 10 ** 8), (30, 400)):
        assert R.draw_expon_dis_many(mean, seeds, total) == [int(R.draw_expon_dis(mean, int(s), total)) for s in seeds]


def test_sampling_fast_path_equals_slow_path():
    """sampling() with the vectorised first-attempt lengths returns exactly the reads of the per-seed path."""
    import random
    from seq2squiggle_b200 import reads as R
    rng = np.random.default_rng(5)
    genome = ["".join(rng.choice(list("ACGTN"), 6000, p=[0.24, 0.24, 0.24, 0.24, 0.04])) for _ in range(2)]
    lens = [len(g) for g in genome]
    random.seed(9)
    fast = R.sampling(300, genome, lens, 400, 9, sum(lens), "expon", "dna-r10-prom")
    random.seed(9)
    orig = R.draw_expon_dis_many
    try:
        R.draw_expon_dis_many = lambda mean, seeds, total: [int(R.draw_expon_dis(mean, int(s), total)) for s in seeds]
        slow_same_seed = R.sampling(300, genome, lens, 400, 9, sum(lens), "expon", "dna-r10-prom")
    finally:
        R.draw_expon_dis_many = orig
    assert fast == slow_same_seed and len(fast) > 250


def test_streamed_sampling_equals_eager(tmp_path):
    """get_reads(stream=True) (what the single-process inference_run uses to overlap sampling with the GPU) yields the
    same reads in the same order as the eager list; cheap names are unique."""
    import random
    from seq2squiggle_b200 import reads as R
    rng = np.random.default_rng(8)
    g = "".join(rng.choice(list("ACGTN"), 30000, p=[0.245, 0.245, 0.245, 0.245, 0.02]))
    fasta = tmp_path / "g.fasta"
    fasta.write_text(">chr1\n" + "\n".join(g[i:i + 70] for i in range(0, len(g), 70)) + "\n")
    cfg = {"max_dna_len": 16}
    random.seed(4)
    eager, hint = R.get_reads(str(fasta), False, 400, 600, -1, cfg, "expon", 4, "dna-r10-prom", 30)
    eager = [s for s, _ in eager]
    random.seed(4)
    lazy, hint2 = R.get_reads(str(fasta), False, 400, 600, -1, cfg, "expon", 4, "dna-r10-prom", 30, stream=True,
                              cheap_names=True)
    lazy = list(lazy)
    assert hint > 0 and hint2 is None
    assert [s for s, _ in lazy] == eager and len(eager) > 350
    assert len({n for _, n in lazy}) == len(lazy)
    with pytest.raises(ValueError):      # argument validation still happens at call time, not at first iteration
        R.get_reads(str(fasta), False, -1, 600, -1, cfg, "expon", 4, "dna-r10-prom", 30, stream=True)


def test_sharded_sampling_equals_eager(tmp_path):
    """get_reads_shard (what a torchrun rank uses): the shards of all ranks, concatenated, are the eager read list —
    although no rank materialises another rank's reads — the ``random`` state after a full replay is the eager one,
    and the first-chunk indices are the running chunk totals.  Genome with N runs and three contigs, DNA and RNA."""
    import random
    from seq2squiggle_b200 import reads as R
    from seq2squiggle_b200.inference import chunks_of_read, shard_reads
    rng = np.random.default_rng(12)
    contigs = ["".join(rng.choice(list("ACGTN"), n, p=[0.24, 0.24, 0.24, 0.24, 0.04])) for n in (9000, 20000, 4000)]
    fasta = tmp_path / "g.fasta"
    fasta.write_text("".join(f">c{i}\n{g}\n" for i, g in enumerate(contigs)))
    cfg = {"max_dna_len": 16, "seq_kmer": 9}
    for profile in ("dna-r10-prom", "rna-004-prom"):
        random.seed(6)
        eager = [s for s, _ in R.get_reads(str(fasta), False, 500, 700, -1, cfg, "expon", 6, profile, 30)[0]]
        after_eager = random.random()
        assert 400 < len(eager) <= 500
        counts = [chunks_of_read(len(s), 9) for s in eager]
        for world in (1, 2, 3, 8):
            got, bases = [], []
            for rank in range(world):
                random.seed(6)
                it, (lo, hi), n_all, base = R.get_reads_shard(str(fasta), False, 500, 700, -1, cfg, "expon", 6, profile,
                                                              30, rank, world, shard_reads, chunks_of_read,
                                                              cheap_names=True)
                mine = list(it)
                assert n_all == len(eager) and len(mine) == hi - lo and base == sum(counts[:lo])
                assert [n for _, n in mine] == [f"read_{i}" for i in range(lo, hi)]
                got += [s for s, _ in mine]
                bases.append(base)
            assert got == eager and bases == sorted(bases)
        # lengths-only replay: same lengths, same generator state afterwards
        random.seed(6)
        seqs, lens = R.preprocess_genome(str(fasta))
        ln = list(R.sampling_iter(500, seqs, lens, 700, 6, sum(lens), "expon", profile, lengths_only=True))
        assert ln == [len(s) for s in eager] and random.random() == after_eager
    # read mode shards the sampled list
    rd = tmp_path / "reads.fasta"
    rd.write_text("".join(f">r{i}\n{'ACGT' * (10 + 7 * i)}\n" for i in range(6)))
    full = [s for s, _ in R.get_reads(str(rd), True, 40, 0, -1, cfg, "expon", 3, "dna-r10-prom", 30)[0]]
    parts = []
    for rank in range(3):
        it, (lo, hi), n_all, base = R.get_reads_shard(str(rd), True, 40, 0, -1, cfg, "expon", 3, "dna-r10-prom", 30, rank,
                                                      3, shard_reads, chunks_of_read)
        parts += [s for s, _ in it]
        assert n_all == 40
    assert parts == full


def test_contig_lookup_by_bisection_equals_reference_walk():
    """sampling_iter finds (contig, offset) of a genome-wide position by bisection; utils.py:359-371 walks the contigs.
    Same answer for every position, including contigs of length 0 and the last base."""
    import itertools
    from bisect import bisect_right
    lens = [5, 0, 1, 7, 0, 0, 3]
    cum = list(itertools.accumulate(lens))
    for pos in range(sum(lens)):
        gi = bisect_right(cum, pos)
        assert (gi, pos - (cum[gi - 1] if gi else 0)) == R.get_genome_and_position(lens, pos)
    with pytest.raises(ValueError):
        R.get_genome_and_position(lens, sum(lens))


def test_fasta_fastq_parser(tmp_path):
    fa = tmp_path / "a.fasta"
    fa.write_text(">r1 desc here\nACGT\nacgtNN\n\n>r2\nTTTT\n>empty\n>r3\tx\nGG\r\nCC\r\n")
    assert list(R.read_fasta(fa)) == [("ACGTacgtNN", "r1"), ("TTTT", "r2"), ("", "empty"), ("GGCC", "r3")]
    fq = tmp_path / "b.fastq"
    fq.write_text("@q1 comment\nACGTAC\n+\n@@@@II\n@q2\nGGA\nTT\n+q2\nII\nIII\n")
    assert list(R.read_fasta(fq)) == [("ACGTAC", "q1"), ("GGATT", "q2")]
    gz = tmp_path / "c.fasta.gz"
    with gzip.open(gz, "wt") as fh:
        fh.write(">g\nAC\nGT\n")
    assert list(R.read_fasta(gz)) == [("ACGT", "g")]


def test_example_files_shapes():
    """SURVEY §8c integer facts about the reference's example reads (only if the reference tree is mounted)."""
    path = "/root/reference/example/lamda_genome_reads.fasta"
    if not os.path.exists(path):
        pytest.skip("reference examples not present on this box")
    lens = [len(s) for s, _ in R.read_fasta(path)]
    assert lens == [2843, 10510, 8487, 2207, 11936, 4407, 1434, 2449, 14971, 11072]


def test_genome_preprocessing_and_revcomp():
    assert R.process_genome("acgtRYn-x") == ("ACGTNNNNN", 9)
    assert R.reverse_complement("AACGTN") == "NACGTT"
    assert R.read_check("ACGT" * 10, 40, 0, "dna-r10-prom", 30)
    assert not R.read_check("ACGT" * 5, 40, 0, "dna-r10-prom", 30)       # cut short by the genome end
    assert R.read_check("ACGT" * 10, 50, 0, "rna-004-prom", 30)          # RNA: length mismatch allowed
    assert not R.read_check("N" * 5 + "A" * 35, 40, 0, "dna-r10-prom", 30)


def test_get_reads_modes_and_errors(tmp_path):
    cfg = {"max_dna_len": 16, "seq_kmer": 9}
    fa = tmp_path / "g.fasta"
    rng = np.random.default_rng(1)
    fa.write_text(">chr\n" + "".join(rng.choice(list("ACGT"), 5000)) + "\n")
    with pytest.raises(ValueError, match="coverage c or the number of reads n"):
        R.get_reads(fa, False, -1, 1000, -1, cfg, "expon", 3, "dna-r10-prom", 30)
    with pytest.raises(ValueError, match="not both"):
        R.get_reads(fa, False, 5, 1000, 3, cfg, "expon", 3, "dna-r10-prom", 30)
    with pytest.raises(ValueError, match="read length r"):
        R.get_reads(fa, False, 5, 0, -1, cfg, "expon", 3, "dna-r10-prom", 30)
    random.seed(3)
    a = [s for s, _ in R.get_reads(fa, False, 25, 300, -1, cfg, "expon", 3, "dna-r10-prom", 30)[0]]
    random.seed(3)
    b = [s for s, _ in R.get_reads(fa, False, 25, 300, -1, cfg, "expon", 3, "dna-r10-prom", 30)[0]]
    assert a == b and 0 < len(a) <= 25 and all(len(s) >= 30 for s in a)
    random.seed(3)
    cov, _ = R.get_reads(fa, False, -1, 500, 2, cfg, "expon", 3, "dna-r10-prom", 30)
    assert len(list(cov)) <= round(2 * 5000 / 500)
    # read mode: every read once, names preserved; with -n: sampled with replacement by random.Random(seed)
    rd = tmp_path / "r.fasta"
    rd.write_text(">a\nACGTACGTACGTAAA\n>b\nGGGGGGGGGGCCCCCCCCCCTTTT\n")
    gen, total = R.get_reads(rd, True, -1, 1000, -1, cfg, "expon", 3, "dna-r10-prom", 30)
    assert list(gen) == [("ACGTACGTACGTAAA", "a"), ("GGGGGGGGGGCCCCCCCCCCTTTT", "b")] and total == 39
    gen, _ = R.get_reads(rd, True, 7, 1000, -1, cfg, "expon", 5, "dna-r10-prom", 30)
    seqs = [s for s, _ in gen]
    rr = random.Random(5)
    assert seqs == [rr.choice(["ACGTACGTACGTAAA", "GGGGGGGGGGCCCCCCCCCCTTTT"]) for _ in range(7)]
