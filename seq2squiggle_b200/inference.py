"""``inference_run`` — the Python API behind ``seq2squiggle predict`` (reference ``inference.py:270-427``).

Same 30 keyword arguments, same derived values (``dwell_mean = sample_rate / bps``, ``ideal_mode``, ``seq_kmer``
from the profile), same writer selection by file extension and the same exceptions.  What differs is below the
plug point: instead of a Lightning ``Trainer.predict`` over a one-hot DataLoader, whole reads are batched by chunk
count and pushed through ``seq2squiggle.predict_reads`` (``model.py`` here), i.e. ``s2s_forward_reads`` with copies
and the file writer overlapped.

Multi-GPU (``torchrun --nproc-per-node N -m seq2squiggle_b200 predict ...``): one process per GPU.  Every rank
derives the same read list from the seed, takes a contiguous range of reads balanced by chunk count
(``shard_reads``), keys its Philox draws by the *global* chunk index (results do not depend on N) and writes
``<stem>.part<rank>.blow5`` with its GLOBAL read numbers / ids.  At the end rank 0's part becomes ``<out>`` and every
other rank splices its own records into it at its byte offset, all ranks in parallel (``splice_blow5_part``: kernel
``copy_file_range`` + a vectorised ``start_time`` shift), instead of one rank re-writing the whole output record by
record (``merge_blow5_parts``, kept as the stand-alone tool).  No collective touches the data path; the only
communication is one small all-gather (bytes and samples per rank) and two barriers on the gloo control plane.
"""
from __future__ import annotations

import logging
import os
import struct
from typing import Iterable, Iterator, List, Sequence, Tuple

import numpy as np

from .checkpoint import check_model
from .profiles import get_profile, update_config, update_profile
from .reads import get_reads, get_reads_shard
from .signal_io import BLOW5Writer, POD5Writer, indexed_uuid

logger = logging.getLogger("seq2squiggle")

BATCH_CHUNKS = int(os.environ.get("S2S_READ_BATCH_CHUNKS", 131072))  # chunks per predict_reads() call


def get_writer(out, profile, ideal_mode, export_every_n_samples, profile_name, preserve_read_ids):
    """inference.py:28-82: writer by extension; an existing output file is deleted."""
    out = str(out)
    out_base = os.path.basename(out)
    out_dir = os.path.dirname(out)
    if out_dir and not os.path.exists(out_dir):
        os.makedirs(out_dir, exist_ok=True)
    if os.path.exists(out):
        logger.warning(f"Output file {out} already exists. File will be deleted.")
        os.remove(out)
    if any(out_base.endswith(ext) for ext in (".blow5", ".slow5")):
        return BLOW5Writer(out, profile, ideal_mode, profile_name, preserve_read_ids), export_every_n_samples
    if out_base.endswith(".pod5"):
        logger.warning("POD5 Writer does not support appending to an existing file.")
        logger.warning("All simulated reads will be stored in RAM before exporting to target pod5.")
        logger.warning("This might lead to Out of Memory errors for large-scale simulations. Consider exporting to "
                       "BLOW5/SLOW5 and using the blue_crab tool for conversion to pod5.")
        return POD5Writer(out, profile, ideal_mode, profile_name, preserve_read_ids), float("inf")
    logger.error("Output file must have .pod5, .slow5, or .blow5 extension.")
    raise ValueError("Output file must have .pod5, .slow5, or .blow5 extension.")


def get_saved_weights(profile_name) -> str:
    """inference.py:85-221 downloads release weights from GitHub; there is no network path here."""
    raise PermissionError("seq2squiggle_b200 does not download model weights. Download compatible weights manually "
                          "from the seq2squiggle GitHub repository "
                          "(https://github.com/ZKI-PH-ImageAnalysis/seq2squiggle) and specify these using the "
                          "`--model` parameter")


# --------------------------------------------------------------------------------------------------
# read batching and sharding (host logic; covered by CPU tests incl. world_size-2 gloo)
# --------------------------------------------------------------------------------------------------
def chunks_of_read(read_len: int, k: int, max_dna: int = 16) -> int:
    n = read_len - k + 1
    return 0 if n <= 0 else -(-n // max_dna)


def batch_reads(reads: Iterable[Tuple[str, str]], k: int, batch_chunks: int = BATCH_CHUNKS) -> Iterator[list]:
    """Groups whole reads into batches of about ``batch_chunks`` chunks (a read is never split)."""
    cur, n = [], 0
    for item in reads:
        cur.append(item)
        n += chunks_of_read(len(item[0]), k)
        if n >= batch_chunks:
            yield cur
            cur, n = [], 0
    if cur:
        yield cur


def shard_reads(chunk_counts: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous read ranges ``[lo, hi)`` per rank, balanced by chunk count: rank r ends at the first read where
    the running chunk total reaches ``(r+1)/world_size`` of the whole."""
    counts = np.asarray(chunk_counts, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        bounds.append(max(int(np.searchsorted(cum, target, side="left")), bounds[-1]))
    bounds.append(len(counts))
    bounds = [min(b, len(counts)) for b in bounds]
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def merge_blow5_parts(out: str, parts: Sequence[str], preserve_read_ids: bool) -> Tuple[int, int]:
    """Stitches uncompressed BLOW5 part files (one per rank, in rank order) into ``out``: the first part's header is
    kept; every record's read_number / start_time (and the synthetic read id) is shifted by the totals of the
    parts before it, which is what a single writer would have produced.  Returns (reads, samples)."""
    reads_total = samples_total = 0
    with open(out, "wb") as fo:
        for pi, part in enumerate(parts):
            with open(part, "rb") as fi:
                head = fi.read(64)
                if head[:6] != b"BLOW5\x01":
                    raise ValueError(f"{part} is not a BLOW5 file")
                if head[9] != 0:
                    raise ValueError("merge_blow5_parts needs uncompressed records")
                (hsize,) = struct.unpack("<I", fi.read(4))
                ascii_hdr = fi.read(hsize)
                if pi == 0:
                    fo.write(head + struct.pack("<I", hsize) + ascii_hdr)
                base_reads, base_samples = reads_total, samples_total
                while True:
                    szb = fi.read(8)
                    if szb[:5] == b"5WOLB" or len(szb) < 8:
                        break
                    (size,) = struct.unpack("<Q", szb)
                    body = bytearray(fi.read(size))
                    (idl,) = struct.unpack_from("<H", body, 0)
                    (siglen,) = struct.unpack_from("<Q", body, 2 + idl + 4 + 32)
                    (rnum,) = struct.unpack_from("<i", body, size - 13)
                    (stime,) = struct.unpack_from("<Q", body, size - 8)
                    struct.pack_into("<i", body, size - 13, rnum + base_reads)
                    struct.pack_into("<Q", body, size - 8, stime + base_samples)
                    if not preserve_read_ids:
                        new_id = str(indexed_uuid(rnum + base_reads + 1)).encode()
                        if len(new_id) == idl:
                            body[2:2 + idl] = new_id
                    fo.write(szb)
                    fo.write(body)
                    reads_total = max(reads_total, rnum + base_reads + 1)
                    samples_total += siglen
        fo.write(b"5WOLB")
    return reads_total, samples_total


def blow5_record_span(path: str) -> Tuple[int, int]:
    """``(first byte of the first record, byte after the last record)`` of an uncompressed BLOW5 file."""
    with open(path, "rb") as f:
        head = f.read(64)
        if head[:6] != b"BLOW5\x01":
            raise ValueError(f"{path} is not a BLOW5 file")
        if head[9] != 0:
            raise ValueError("splicing BLOW5 parts needs uncompressed records")
        (hsize,) = struct.unpack("<I", f.read(4))
        f.seek(0, os.SEEK_END)
        end = f.tell()
        f.seek(end - 5)
        if f.read(5) != b"5WOLB":
            raise ValueError(f"{path} has no end-of-file marker (writer not closed?)")
    return 64 + 4 + hsize, end - 5


def splice_blow5_part(out: str, part: str, dst_offset: int, start_time_shift: int) -> int:
    """Copies the records of ``part`` into ``out`` at byte ``dst_offset`` and adds ``start_time_shift`` (the samples
    written by the ranks before this one) to every record's ``start_time`` — the last eight bytes of a record, as in
    ``merge_blow5_parts``.  Run by every rank > 0 at the same time on disjoint byte ranges of ``out``.  The bulk copy is
    ``os.copy_file_range`` (in-kernel; falls back to read + pwrite), the shift one gather / add / scatter over a memory
    map of the part file (which is modified).  Returns the number of records."""
    lo, hi = blow5_record_span(part)
    n_bytes = hi - lo
    if n_bytes == 0:
        return 0
    src = os.open(part, os.O_RDONLY)
    try:
        ends, pos = [], lo                                   # record boundaries from the 8-byte size prefixes
        while pos < hi:
            (size,) = struct.unpack("<Q", os.pread(src, 8, pos))
            pos += 8 + size
            ends.append(pos - lo)
    finally:
        os.close(src)
    if pos != hi:
        raise ValueError(f"{part}: records do not end at the end-of-file marker")
    if start_time_shift:     # patched in the rank's own part file, before the copy: no page of `out` is shared between ranks
        mm = np.memmap(part, dtype=np.uint8, mode="r+", offset=lo, shape=(n_bytes,))
        idx = (np.asarray(ends, dtype=np.int64) - 8)[:, None] + np.arange(8, dtype=np.int64)[None, :]
        st = np.ascontiguousarray(mm[idx]).view("<u8")[:, 0] + np.uint64(start_time_shift)
        mm[idx] = st.astype("<u8").view(np.uint8).reshape(-1, 8)
        mm.flush()
        del mm
    src = os.open(part, os.O_RDONLY)
    dst = os.open(out, os.O_RDWR)
    try:
        done = 0
        use_cfr = hasattr(os, "copy_file_range")
        while done < n_bytes:
            want = min(n_bytes - done, 1 << 30)
            if use_cfr:
                try:
                    got = os.copy_file_range(src, dst, want, lo + done, dst_offset + done)
                    if got <= 0:
                        raise OSError("copy_file_range copied nothing")
                    done += got
                    continue
                except OSError:
                    use_cfr = False                          # e.g. across file systems on an old kernel
            buf = os.pread(src, min(want, 64 << 20), lo + done)
            os.pwrite(dst, buf, dst_offset + done)
            done += len(buf)
    finally:
        os.close(src)
        os.close(dst)
    return len(ends)


def splice_parts_collective(out: str, my_part: str, rank: int, world: int, my_samples: int, dist) -> Tuple[int, int]:
    """The end of a sharded run, called by every rank once its writer has closed ``my_part``: rank 0's part becomes
    ``out`` (a rename), the others splice their records in at their byte offsets at the same time.  The parts carry
    global read numbers / ids already (``writer._id_base`` = the shard's first read); only ``start_time``, a running sum
    over all earlier reads of the run, needs the totals of the earlier ranks.  Returns (bytes, samples) of ``out``."""
    lo, hi = blow5_record_span(my_part)
    info = [None] * world
    dist.all_gather_object(info, (hi - lo, int(my_samples), lo))
    body = [b for b, _, _ in info]
    samples = [n for _, n, _ in info]
    first = info[0][2]                                       # the kept header is rank 0's
    total = first + sum(body) + 5
    if rank == 0:
        os.replace(my_part, out)
        with open(out, "r+b") as f:
            f.truncate(total)                                # drops rank 0's end marker / reserves the other ranks' ranges
    dist.barrier()
    if rank > 0:
        splice_blow5_part(out, my_part, first + sum(body[:rank]), sum(samples[:rank]))
        os.remove(my_part)
    dist.barrier()
    if rank == 0:
        with open(out, "r+b") as f:
            f.seek(total - 5)
            f.write(b"5WOLB")
    return total, sum(samples)


def part_path(out: str, rank: int) -> str:
    """Per-rank part file of a multi-GPU run: ``sim.blow5`` -> ``sim.part<rank>.blow5`` (keeps the extension, so the
    writer factory's extension check applies to it as to any output).  ``S2S_PART_DIR`` moves the parts of the ranks
    > 0 to another directory — a tmpfs such as /dev/shm turns "write the part, then copy it into the output" into one
    pass over the disk; rank 0's part stays beside ``out`` because it *becomes* the output by a rename."""
    stem, ext = os.path.splitext(str(out))
    part_dir = os.environ.get("S2S_PART_DIR")
    if part_dir and rank > 0:
        stem = os.path.join(part_dir, os.path.basename(stem))
    return f"{stem}.part{rank}{ext}"


def _dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


# --------------------------------------------------------------------------------------------------
def inference_run(config: dict, saved_weights: str, fasta: str, read_input: bool, n: int, r: int, c: int, out: str,
                  profile: str, dwell_mean, dwell_std: float, noise_std: float, noise_sampling: bool,
                  duration_sampling: bool, distr: str, predict_batch_size: int, export_every_n_samples: int,
                  sample_rate, bps, digitisation, range_val, offset_mean, offset_std, median_before_mean,
                  median_before_std, min_noise: float, min_duration: float, min_read_len: int,
                  preserve_read_ids: bool, seed: int, precision: str = "fp16"):
    """inference.py:270-427.  ``predict_batch_size`` is accepted for compatibility (the engine sizes its own
    sub-batches); ``precision`` ("fp16" tensor-core path, or "fp32" parity path) is the one added argument."""
    import torch
    from .model import seq2squiggle

    profile_dict = get_profile(profile)
    profile_dict = update_profile(profile_dict, sample_rate=sample_rate, bps=bps, digitisation=digitisation,
                                  range=range_val, offset_mean=offset_mean, offset_std=offset_std,
                                  median_before_mean=median_before_mean, median_before_std=median_before_std)
    if dwell_mean is None:
        dwell_mean = profile_dict["sample_rate"] / profile_dict["bps"]
    config = update_config(profile, config)
    ideal_mode = not (duration_sampling or dwell_std > 0)

    rank, world, local = _dist_env()
    out = str(out)
    my_out = out
    if world > 1:
        if not out.endswith(".blow5"):
            raise ValueError("multi-GPU predict writes BLOW5 part files: use a .blow5 output")
        my_out = part_path(out, rank)
        if rank == 0 and os.path.exists(out):
            logger.warning(f"Output file {out} already exists. File will be deleted.")
            os.remove(out)
    writer, export_every_n_samples = get_writer(my_out if world > 1 else out, profile_dict, ideal_mode,
                                                export_every_n_samples, profile_name=profile,
                                                preserve_read_ids=preserve_read_ids)
    if world > 1:
        writer.filename = my_out
    if saved_weights is None:
        saved_weights = get_saved_weights(profile)

    local = local % max(torch.cuda.device_count(), 1)   # more ranks than GPUs (tests): ranks share devices
    torch.cuda.set_device(local)
    load_model = seq2squiggle.load_from_checkpoint(
        checkpoint_path=saved_weights, out_writer=writer, dwell_mean=dwell_mean, dwell_std=dwell_std,
        noise_std=noise_std, noise_sampling=noise_sampling, duration_sampling=duration_sampling,
        export_every_n_samples=export_every_n_samples, min_noise=min_noise, min_duration=min_duration, device=local,
        precision=precision)
    check_model(load_model.hparams.config, config)

    # reads are sampled lazily, batch by batch, while the GPU works on the previous batches (the sampler is
    # sequential Python); a sharded multi-process run first replays the sampler for the read lengths alone to
    # balance the ranks by chunk count, then materialises only its own reads (reads.get_reads_shard)
    k = config["seq_kmer"]
    chunk_base = 0
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("gloo")      # control plane only: a barrier before the merge
        reads, (lo, hi), n_all, chunk_base = get_reads_shard(
            fasta, read_input, n, r, c, config, distr, seed, profile, min_read_len, rank, world, shard_reads,
            chunks_of_read, cheap_names=not preserve_read_ids)
        logger.info(f"rank {rank}/{world}: reads [{lo}, {hi}) of {n_all}, first global chunk {chunk_base}")
        writer._id_base = lo                       # read_number / synthetic read ids are global from the start
        np.random.seed((seed + rank) % (2 ** 32))  # per-record offset / median_before draws differ per rank
    else:
        reads, total_l = get_reads(fasta, read_input, n, r, c, config, distr, seed, profile, min_read_len,
                                   stream=True, cheap_names=not preserve_read_ids)
    load_model.chunks_done = chunk_base
    n_reads = 0
    for batch in batch_reads(reads, k):
        load_model.predict_reads(batch)
        n_reads += len(batch)
    load_model.on_predict_epoch_end()
    stats = getattr(load_model, "last_stats", None)
    logger.info(f"rank {rank}: simulated {n_reads} reads, {writer.samples_written} samples -> {writer.filename}")

    if world > 1:
        import torch.distributed as dist
        if not os.path.exists(my_out):           # a rank without reads still contributes an (empty) part
            writer.signals = {}
            writer.save()
        nbytes, ns = splice_parts_collective(out, my_out, rank, world, writer.samples_written, dist)
        if rank == 0:
            logger.info(f"spliced {world} parts: {ns} samples, {nbytes} bytes -> {out}")
        dist.barrier()
    return stats
