"""``inference_run`` — the Python API behind ``seq2squiggle predict`` (reference ``inference.py:270-427``).

Same 30 keyword arguments, same derived values (``dwell_mean = sample_rate / bps``, ``ideal_mode``, ``seq_kmer``
from the profile), same writer selection by file extension and the same exceptions.  What differs is below the
plug point: instead of a Lightning ``Trainer.predict`` over a one-hot DataLoader, whole reads are batched by chunk
count and pushed through ``seq2squiggle.predict_reads`` (``model.py`` here), i.e. ``s2s_forward_reads`` with copies
and the file writer overlapped.

Multi-GPU (``torchrun --nproc-per-node N -m seq2squiggle_b200 predict ...``): one process per GPU.  Every rank
derives the same read list from the seed, takes a contiguous range of reads balanced by chunk count
(``shard_reads``), keys its Philox draws by the *global* chunk index (results do not depend on N) and writes
``<stem>.part<rank>.blow5``; rank 0 stitches the parts into ``<out>`` (``merge_blow5_parts``).  No collective touches the
data path; the only communication is a barrier on the control-plane process group.
"""
from __future__ import annotations

import logging
import os
import struct
import uuid
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


def part_path(out: str, rank: int) -> str:
    """Per-rank part file of a multi-GPU run: ``sim.blow5`` -> ``sim.part<rank>.blow5`` (keeps the extension, so the
    writer factory's extension check applies to it as to any output)."""
    stem, ext = os.path.splitext(str(out))
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
        dist.barrier()
        if rank == 0:
            parts = [part_path(out, i) for i in range(world)]
            nr, ns = merge_blow5_parts(out, parts, preserve_read_ids)
            for p in parts:
                os.remove(p)
            logger.info(f"merged {world} parts: {nr} reads, {ns} samples -> {out}")
        dist.barrier()
    return stats
