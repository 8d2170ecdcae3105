"""Read source of ``seq2squiggle predict``: FASTA/FASTQ parsing, read mode and reference-mode sampling.

Mirrors ``utils.py:290-671`` of the reference (``get_reads`` and everything under it).  The reference leans on
``pysam.FastxFile`` and one ``scipy.stats`` ``rvs`` call per read; neither is needed here:

* ``read_fasta`` is a small FASTA/FASTQ (optionally gzip) parser with FastxFile's semantics (name = first word of
  the header line, multi-line FASTA sequences are joined, sequence returned verbatim — no upper-casing);
* the read-length draws use ``numpy.random.RandomState(seed)`` directly, which is what ``scipy.stats.*.rvs(...,
  random_state=<int>)`` does internally, so a given ``--seed`` yields the *same reads* as the reference
  (pinned by ``tests/test_reads.py`` against scipy and against ``tests/golden/read_sampling.json``, which was
  produced by the reference's own ``utils.sampling``).

The Python ``random`` call order of ``utils.py:415-479`` (start position, strand, N replacement) is kept, since the
module-level generator is seeded once by ``set_seeds`` (``utils.py:722-741``).
"""
from __future__ import annotations

import gzip
import itertools
import logging
import os
import random
import re
from bisect import bisect_right
from typing import Generator, Iterable, List, Sequence, Tuple
from uuid import uuid4

import numpy as np

logger = logging.getLogger("seq2squiggle")


# --------------------------------------------------------------------------------------------------
# FASTA / FASTQ
# --------------------------------------------------------------------------------------------------
def _open_text(path):
    path = str(path)
    with open(path, "rb") as fh:
        magic = fh.read(2)
    if magic == b"\x1f\x8b":
        return gzip.open(path, "rt", encoding="latin-1", newline=None)
    return open(path, "r", encoding="latin-1", newline=None)


def read_fasta(path, rna: bool = False) -> Generator[Tuple[str, str], None, None]:
    """utils.py:290-308: yields ``(sequence, name)`` for every FASTA or FASTQ record of ``path``."""
    with _open_text(path) as fh:
        name, parts = None, []
        line = fh.readline()
        while line:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if name is not None:
                    yield "".join(parts), name
                name, parts = (line[1:].split() or [""])[0], []
                line = fh.readline()
            elif line.startswith("@") and name is None:
                # FASTQ record: header, sequence line(s) up to '+', then as many quality characters
                qname = (line[1:].split() or [""])[0]
                seq_parts = []
                line = fh.readline()
                while line and not line.startswith("+"):
                    seq_parts.append(line.strip())
                    line = fh.readline()
                seq = "".join(seq_parts)
                qlen = 0
                line = fh.readline()
                while line and qlen < len(seq):
                    qlen += len(line.rstrip("\r\n"))
                    line = fh.readline()
                yield seq, qname
            else:
                if name is not None:
                    parts.append(line.strip())
                line = fh.readline()
        if name is not None:
            yield "".join(parts), name


def load_genome(fasta) -> Generator[str, None, None]:
    """utils.py:586-589."""
    for seq, _ in read_fasta(fasta):
        yield str(seq)


def process_genome(genome_seq: str) -> Tuple[str, int]:
    """utils.py:592-595: upper-case, everything outside ACGT becomes N."""
    genome_seq = re.sub(r"[^ATCG]", "N", genome_seq.upper())
    return genome_seq, len(genome_seq)


def preprocess_genome(fasta):
    """utils.py:608-638 (the reference maps over a process pool; the result is the same)."""
    logger.debug("Preprocessing the genome")
    results = [process_genome(s) for s in load_genome(fasta)]
    if not results:
        raise ValueError(f"No sequences found in {fasta}")
    seqs, lens = zip(*results)
    logger.debug("Preprocessing the genome finished.")
    return seqs, lens


def compute_totals(generator) -> Tuple[int, int]:
    """utils.py:598-605."""
    total_reads = total_length = 0
    for sequence, _ in generator:
        total_reads += 1
        total_length += len(sequence)
    return total_reads, total_length


# --------------------------------------------------------------------------------------------------
# read-length laws (utils.py:311-331).  scipy's rvs(size=1, random_state=int) == RandomState(int) draw below.
# --------------------------------------------------------------------------------------------------
def draw_gamma_dis(mean, seed, total_len):
    x = np.random.RandomState(seed).standard_gamma(6.3693711, size=1) + 0.53834893      # st.gamma.rvs(a, loc)
    sample = int((x * mean / 4.39)[0])
    return np.clip(sample, 1, total_len)


def draw_beta_dis(mean, seed, total_len):
    x = np.random.RandomState(seed).beta(1.778, 7.892, size=1) * 34191.257 + 316.758    # st.beta.rvs(a, b, loc, scale)
    sample = (x[0] * mean / 6615.0).astype(int)
    return np.clip(sample, 1, total_len)


def draw_expon_dis(mean, seed, total_len):
    x = np.random.RandomState(seed).standard_exponential(size=1) * 6972.5319847131141 + 213.98910256668592
    sample = (x[0] * mean / 7106.0).astype(int)
    return np.clip(sample, 1, total_len)


DISTR_FUNCS = {"beta": draw_beta_dis, "gamma": draw_gamma_dis, "expon": draw_expon_dis}


def mt19937_first_doubles(seeds: np.ndarray) -> np.ndarray:
    """First ``random_sample()`` of ``numpy.random.RandomState(seed)`` for MANY integer seeds (< 2**32) at once.

    ``RandomState(int)`` seeds MT19937 with Knuth's ``init_genrand`` recurrence and the first double is built from the
    first two tempered 32-bit outputs, ``((y0 >> 5) * 2**26 + (y1 >> 6)) / 2**53``.  Only state words 0, 1, 2, 397 and
    398 of the seeded state enter those two outputs, so the recurrence is run (vectorised over the seeds) up to word
    398 and the two twists are done by hand.  One ``RandomState`` construction per read (~180 us) was 90 % of the
    reference-mode start-up time of ``seq2squiggle predict``; this is bit-identical (tests/test_reads.py) and ~100x
    faster."""
    mt = np.asarray(seeds, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
    keep = {0: mt}
    mask = np.uint64(0xFFFFFFFF)
    for i in range(1, 399):
        mt = (np.uint64(1812433253) * (mt ^ (mt >> np.uint64(30))) + np.uint64(i)) & mask
        if i in (1, 2, 397, 398):
            keep[i] = mt

    def out(lo, hi, far):
        y = (lo & np.uint64(0x80000000)) | (hi & np.uint64(0x7FFFFFFF))
        v = far ^ (y >> np.uint64(1)) ^ np.where(y & np.uint64(1), np.uint64(0x9908B0DF), np.uint64(0))
        v ^= v >> np.uint64(11)
        v ^= (v << np.uint64(7)) & np.uint64(0x9D2C5680)
        v ^= (v << np.uint64(15)) & np.uint64(0xEFC60000)
        v ^= v >> np.uint64(18)
        return v & mask

    a = out(keep[0], keep[1], keep[397]) >> np.uint64(5)
    b = out(keep[1], keep[2], keep[398]) >> np.uint64(6)
    return (a.astype(np.float64) * 67108864.0 + b.astype(np.float64)) / 9007199254740992.0


def draw_expon_dis_many(mean, seeds, total_len) -> List[int]:
    """``[draw_expon_dis(mean, s, total_len) for s in seeds]`` without one RandomState per seed: the legacy
    ``standard_exponential`` is ``-log(1 - random_sample())`` (libm ``log``, hence ``math.log`` per element)."""
    import math
    u = mt19937_first_doubles(np.asarray(seeds, dtype=np.uint64))
    out = []
    for ui in u.tolist():
        x = -math.log(1.0 - ui) * 6972.5319847131141 + 213.98910256668592
        out.append(min(max(int(x * mean / 7106.0), 1), total_len))
    return out


def get_genome_and_position(genome_lengths: Sequence[int], random_position: int) -> Tuple[int, int]:
    """utils.py:359-371."""
    total_length = sum(genome_lengths)
    if random_position >= total_length:
        raise ValueError("Random position exceeds the total length of genomes")
    cumulative = 0
    for i, length in enumerate(genome_lengths):
        cumulative += length
        if random_position < cumulative:
            return i, random_position - (cumulative - length)
    raise ValueError("Random position exceeds the total length of genomes")


def read_check(read: str, read_length: int, read_i: int, profile: str, min_read_len: int = 30) -> bool:
    """utils.py:381-398."""
    if profile.startswith("dna") and len(read) != read_length:
        logger.debug(f"Sampled Read length ({len(read)}) of read {read_i} is shorter than real read length ({read_length}).")
        return False
    if len(read) < min_read_len:
        logger.debug(f"Sampled Read length ({len(read)}) of read {read_i} is shorter than the minimal read length ({min_read_len}).")
        return False
    count_n = read.count("N")
    if count_n > 0.1 * read_length:
        logger.debug(f"Too many 'N' bases ({count_n} out of {read_length}) for read {read_i}")
        return False
    return True


def N_to_ACTG(read: str) -> str:
    """utils.py:401-402 (one ``random.choice`` per N, in order)."""
    return "".join(random.choice("ACGT") if base == "N" else base for base in read)


_COMPLEMENT = str.maketrans("ATCG", "TAGC")


def reverse_complement(f: str) -> str:
    """utils.py:409-412: A<->T, C<->G, everything else unchanged."""
    return f.translate(_COMPLEMENT)[::-1]


def sampling(num_seqs, genome_seqs, genome_lens, r, seed, total_len, distr, profile, min_read_len=30,
             max_retries=20) -> List[str]:
    """utils.py:415-479: sample ``num_seqs`` reads from the genome(s)."""
    return list(sampling_iter(num_seqs, genome_seqs, genome_lens, r, seed, total_len, distr, profile, min_read_len,
                              max_retries))


def sampling_iter(num_seqs, genome_seqs, genome_lens, r, seed, total_len, distr, profile, min_read_len=30,
                  max_retries=20, window=None, lengths_only=False) -> Generator:
    """``sampling`` as a generator: the same reads in the same order (the same calls on the ``random`` module and the
    same per-read seeds), produced on demand so that the caller can overlap sampling with the GPU.

    ``window=(lo, hi)`` — or a sorted list of such disjoint ranges (a rank's batches of a round-robin sharded run) —
    yields only the accepted reads number ``lo <= i < hi`` and stops after the last ``hi``; ``lengths_only`` yields the
    length of every accepted read instead of the read.  Reads that are not
    materialised still consume exactly the ``random`` calls the reference would make for them (start position,
    strand, one ``choice`` per ``N``), so every rank of a sharded run sees the same read list without building it:
    the acceptance test of ``read_check`` is evaluated on (start, length) and ``str.count`` over the genome span
    instead of on a slice."""
    draw = DISTR_FUNCS[distr]
    total_genome_len = sum(genome_lens)
    dna = profile.startswith("dna")
    if window is None:
        windows = [(0, None)]
    elif len(window) == 2 and not isinstance(window[0], (tuple, list)):
        windows = [tuple(window)]
    else:
        windows = [tuple(w) for w in window if w[1] > w[0]] or [(0, 0)]
    wi = 0
    lo, hi = windows[0]
    last_hi = windows[-1][1]
    # first-attempt lengths of all reads in one vectorised pass (default law, 32-bit seeds); retries and the other
    # laws take the per-seed path.  Identical values either way.
    # (computed block by block on demand: a block's arrays stay in cache and the first read is not held up by the rest)
    vectorised = distr == "expon" and r > 0 and num_seqs > 0 and 0 <= seed and seed + num_seqs * (max_retries + 1) < 2 ** 32
    first_len, block0, kBlock = None, 0, 8192
    has_n = ["N" in g for g in genome_seqs]
    cum_lens = list(itertools.accumulate(genome_lens))
    debug = logger.isEnabledFor(logging.DEBUG)
    accepted = 0
    randint, choice = random.randint, random.choice
    for read_i in range(num_seqs):
        if last_hi is not None and accepted >= last_hi:
            return
        while hi is not None and accepted >= hi and wi + 1 < len(windows):   # next window of this rank
            wi += 1
            lo, hi = windows[wi]
        retries = 0
        while retries < max_retries:
            start_pos = randint(0, total_genome_len - 1)
            # get_genome_and_position (utils.py:359-371) by bisection: a transcriptome has 10^5 sequences, and the
            # reference's linear walk (plus a sum over all lengths) per read is what dominates there
            genome_index = bisect_right(cum_lens, start_pos)
            start_index = start_pos - (cum_lens[genome_index - 1] if genome_index else 0)
            genome = genome_seqs[genome_index]
            if vectorised and retries == 0:
                if first_len is None or read_i >= block0 + kBlock:
                    block0 = read_i - read_i % kBlock
                    idx = np.arange(block0, min(block0 + kBlock, num_seqs), dtype=np.uint64)
                    first_len = draw_expon_dis_many(r, np.uint64(seed) + idx * np.uint64(max_retries + 1), total_len)
                read_length = first_len[read_i - block0]
            else:
                unique_seed = seed + read_i * (max_retries + 1) + retries
                read_length = int(draw(r, unique_seed, total_len)) if r > 0 else len(genome)
            read_strand = choice("+-") if dna else "+"
            # read_check (utils.py:381-398) on the span [start_index, start_index + read_length) of the genome
            got = max(0, min(read_length, len(genome) - start_index))
            count_n = genome.count("N", start_index, start_index + got) if has_n[genome_index] else 0
            ok = True
            if dna and got != read_length:
                ok = False
                if debug:
                    logger.debug(f"Sampled Read length ({got}) of read {read_i} is shorter than real read length ({read_length}).")
            elif got < min_read_len:
                ok = False
                if debug:
                    logger.debug(f"Sampled Read length ({got}) of read {read_i} is shorter than the minimal read length ({min_read_len}).")
            elif count_n > 0.1 * read_length:
                ok = False
                if debug:
                    logger.debug(f"Too many 'N' bases ({count_n} out of {read_length}) for read {read_i}")
            if ok:
                wanted = accepted >= lo and (hi is None or accepted < hi) and not lengths_only
                if wanted:
                    read = genome[start_index:start_index + got]
                    if count_n:
                        read = N_to_ACTG(read)
                    if read_strand == "-":
                        read = reverse_complement(read)
                    yield read
                else:
                    for _ in range(count_n):      # N_to_ACTG draws one base per N, in order
                        choice("ACGT")
                    if lengths_only:
                        yield got
                accepted += 1
                break
            retries += 1
            if debug:
                if retries >= max_retries:
                    logger.debug(f"Failed to sample a valid read after {max_retries} retries for read {read_i}. Skipping this read.")
                else:
                    logger.debug(f"Retrying to sample read {read_i} (attempt {retries + 1}/{max_retries})")


def export_fasta(read_l: Iterable[str], fasta) -> str:
    """utils.py:480-486 (the reference writes the uuid without a '>' prefix; kept)."""
    file_name, _ = os.path.splitext(str(fasta))
    out_file = f"{file_name}_reads.fasta"
    with open(out_file, "w") as f:
        for read in read_l:
            f.write(f"{str(uuid4())}\n{''.join(read)}\n")
    return out_file


def yield_reads(reads: Iterable[str], cheap_names: bool = False, first=0):
    """utils.py:489-490: ``(read, uuid4 name)``.  ``cheap_names``: the names are only dictionary keys (the writers
    replace them with indexed ids unless ``--preserve-read-ids``), so a counter (starting at ``first``, the global
    index of a shard's first read; or the global indices themselves as a list of ``(lo, hi)`` ranges) does instead of
    100k ``uuid4()`` calls."""
    if cheap_names:
        if isinstance(first, (list, tuple)):
            idx = itertools.chain.from_iterable(range(lo, hi) for lo, hi in first)
            return ((read, f"read_{i}") for i, read in zip(idx, reads))
        return ((read, f"read_{i}") for i, read in enumerate(reads, first))
    return ((read, str(uuid4())) for read in reads)


def sample_reads_from_reference(genome_seqs, genome_lens, n, r, c, config, fasta, seed, save=False, distr="expon",
                                profile="dna-r10-min", min_read_len=30, stream=False, cheap_names=False, window=None,
                                lengths_only=False):
    """utils.py:493-582: argument validation (same messages) + sampling.  ``stream``: return a lazy generator (and no
    length hint) instead of sampling every read up front; with ``window`` / ``lengths_only`` (see ``sampling_iter``)
    the generator yields one shard of the reads / the accepted read lengths."""
    logger.debug("Generating reads from the reference input file.")
    if n <= 0 and c <= 0:
        logger.error("You need to specify the coverage c or the number of reads n")
        raise ValueError("You need to specify the coverage c or the number of reads n")
    if n != -1 and c != -1:
        logger.error("You can only either specify the coverage c or the number of reads, but not both")
        raise ValueError("You can only either specify the coverage c or the number of reads, but not both")
    if r <= 0:
        logger.error("You need to specify an average read length r for sampling from the reads from the reference sequence.")
        raise ValueError("You need to specify the read length r")
    total_len = sum(len(seq) for seq in genome_seqs)
    avg_genome_len = total_len / len(genome_seqs)
    seq_num = n if n != -1 else round(c * total_len / r)
    logger.debug(f"Number of reads: {seq_num}")
    if r > avg_genome_len and profile.startswith("dna"):
        logger.warning(
            f"Average reference sequence length ({avg_genome_len:.2f}) is smaller than the desired average read length ({r})."
            " If the sampled read length is higher than the reference sequence length, they will be skipped."
            " Consider reducing the desired average read length via -r.")
    if stream and not save:
        it = sampling_iter(seq_num, genome_seqs, genome_lens, r, seed, total_len, distr, profile, min_read_len,
                           window=window, lengths_only=lengths_only)
        if lengths_only:
            return it, None
        if window and isinstance(window[0], (tuple, list)):
            return yield_reads(it, cheap_names, first=list(window)), None
        return yield_reads(it, cheap_names, first=window[0] if window else 0), None
    read_list = sampling(seq_num, genome_seqs, genome_lens, r, seed, total_len, distr, profile, min_read_len)
    total_l = sum(round(len(read) / config["max_dna_len"]) for read in read_list)
    reads_fasta = export_fasta(read_list, fasta) if save else yield_reads(read_list, cheap_names)
    logger.debug("Generating reads finished.")
    return reads_fasta, total_l


def get_reads(fasta, read_input, n, r, c, config, distr, seed, profile, min_read_len, save=False, stream=False,
              cheap_names=False):
    """utils.py:641-671: ``(generator of (sequence, name), length hint)``.  ``stream`` / ``cheap_names`` (reference mode
    only): sample lazily (length hint None) / name the reads with a counter instead of ``uuid4()``."""
    logger.info(f"{'Read' if read_input else 'Reference'} mode.")
    is_rna = profile.startswith("rna")
    if read_input:
        if n <= 0:  # every read exactly once
            return read_fasta(fasta, is_rna), compute_totals(read_fasta(fasta, is_rna))[1]
        all_reads = list(read_fasta(fasta, is_rna))
        rng = random.Random(seed)
        sampled = [rng.choice(all_reads) for _ in range(n)]

        def generator():
            for seq, _ in sampled:
                yield seq, str(uuid4())

        return generator(), sum(round(len(seq) / config["max_dna_len"]) for seq, _ in sampled)
    genome_seqs, genome_lens = preprocess_genome(fasta)
    reads_fasta, total_l = sample_reads_from_reference(genome_seqs, genome_lens, n, r, c, config, fasta, seed, save,
                                                       distr, profile, min_read_len, stream, cheap_names)
    return read_fasta(reads_fasta, is_rna) if save else (reads_fasta, total_l)


def get_reads_batches(fasta, read_input, n, r, c, config, distr, seed, profile, min_read_len, rank, world, plan_fn,
                      chunks_fn, cheap_names=False):
    """The reads of rank ``rank`` of a ``world``-process run whose BATCHES are dealt round-robin: ``(iterator of
    (sequence, name) over the rank's batches in order, plan, chunk counts of all reads)`` with ``plan = plan_fn(counts)``
    the list of ``(lo, hi)`` read ranges of all batches; batch ``b`` belongs to rank ``b % world``.  Every rank derives
    the same read list and the same plan from the seed (lengths-only replay of the sampler, about 3 us per read) and
    materialises only its own batches, lazily, while its GPU works."""
    k = config["seq_kmer"]
    if read_input:
        reads, _ = get_reads(fasta, read_input, n, r, c, config, distr, seed, profile, min_read_len)
        reads = list(reads)     # read mode samples references to the input reads: nothing to save
        counts = np.asarray([chunks_fn(len(s), k) for s, _ in reads], dtype=np.int64)
        plan = plan_fn(counts)
        mine = [plan[b] for b in range(rank, len(plan), world)]
        return itertools.chain.from_iterable(reads[lo:hi] for lo, hi in mine), plan, counts
    logger.info("Reference mode.")
    genome_seqs, genome_lens = preprocess_genome(fasta)
    state = random.getstate()
    lens, _ = sample_reads_from_reference(genome_seqs, genome_lens, n, r, c, config, fasta, seed, False, distr, profile,
                                          min_read_len, stream=True, lengths_only=True)
    lens = np.fromiter(lens, dtype=np.int64)
    random.setstate(state)
    nk = lens - k + 1
    counts = np.where(nk > 0, -(-nk // config["max_dna_len"]), 0)
    plan = plan_fn(counts)
    mine = [plan[b] for b in range(rank, len(plan), world)]
    if not mine:
        return iter(()), plan, counts
    reads, _ = sample_reads_from_reference(genome_seqs, genome_lens, n, r, c, config, fasta, seed, False, distr,
                                           profile, min_read_len, stream=True, cheap_names=cheap_names, window=mine)
    return reads, plan, counts


def get_reads_shard(fasta, read_input, n, r, c, config, distr, seed, profile, min_read_len, rank, world, shard_fn,
                    chunks_fn, cheap_names=False):
    """The reads of rank ``rank`` of a ``world``-process run: ``(iterator of (sequence, name), (lo, hi), number of
    reads of the whole run, global index of the shard's first chunk)``.

    Every rank derives the same read list from the seed (the reference has no sharded predict).  In
    reference mode the list is never built: a first pass replays the sampler for the accepted read *lengths* only
    (about 1 us per read), ``shard_fn(chunk counts, world)`` balances the contiguous read ranges by chunk count, and
    a second pass from the same ``random`` state materialises only this rank's reads, lazily, while the GPU works.
    Sampling every read on every rank (16 us and 5 kB per read at ``-r 5000``) cost 10 s and 3 GB per rank at
    BASELINE config 5 (600,000 reads) before the first kernel could start."""
    k = config["seq_kmer"]
    if read_input:
        reads, _ = get_reads(fasta, read_input, n, r, c, config, distr, seed, profile, min_read_len)
        reads = list(reads)     # read mode samples references to the input reads: nothing to save
        counts = np.asarray([chunks_fn(len(s), k) for s, _ in reads], dtype=np.int64)
        lo, hi = shard_fn(counts, world)[rank]
        return iter(reads[lo:hi]), (lo, hi), len(reads), int(counts[:lo].sum())
    logger.info("Reference mode.")
    genome_seqs, genome_lens = preprocess_genome(fasta)
    state = random.getstate()
    lens, _ = sample_reads_from_reference(genome_seqs, genome_lens, n, r, c, config, fasta, seed, False, distr, profile,
                                          min_read_len, stream=True, lengths_only=True)
    lens = np.fromiter(lens, dtype=np.int64)
    random.setstate(state)
    nk = lens - k + 1
    counts = np.where(nk > 0, -(-nk // config["max_dna_len"]), 0)
    lo, hi = shard_fn(counts, world)[rank]
    reads, _ = sample_reads_from_reference(genome_seqs, genome_lens, n, r, c, config, fasta, seed, False, distr,
                                           profile, min_read_len, stream=True, cheap_names=cheap_names,
                                           window=(lo, hi))
    return reads, (lo, hi), len(lens), int(counts[:lo].sum())
