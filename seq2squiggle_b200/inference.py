"""``inference_run`` — the Python API behind ``seq2squiggle predict`` (reference ``inference.py:270-427``).

Same 30 keyword arguments, same derived values (``dwell_mean = sample_rate / bps``, ``ideal_mode``, ``seq_kmer``
from the profile), same writer selection by file extension and the same exceptions.  What differs is below the
plug point: instead of a Lightning ``Trainer.predict`` over a one-hot DataLoader, whole reads are batched by chunk
count and pushed through ``seq2squiggle.predict_reads`` (``model.py`` here), i.e. ``s2s_forward_reads`` with copies
and the file writer overlapped.

Multi-GPU (``torchrun --nproc-per-node N -m seq2squiggle_b200 predict ...``): one process per GPU.  Every rank
derives the same read list from the seed WITHOUT building it (a lengths-only replay of the sampler), cuts it into
batches of about one pipeline piece (``plan_batches``), takes every N-th batch, keys its Philox draws by the *global*
chunk index (results do not depend on N) and writes its batches' records — global read numbers / ids, ``start_time`` and
the per-record NumPy draws of the single-process stream — straight into the ONE output file at the byte offset the owner
of the previous batch publishes (``signal_io.SharedOrder``, a small shared table).  Encoding and ``pwrite`` of the ranks
run in parallel and overlap the compute; there are no part files and no merge pass.  No collective touches the data
path; the gloo control plane carries two barriers (and the random seed of ``-s 0``).
"""
from __future__ import annotations

import logging
import os
from typing import Optional, Iterable, Iterator, List, Sequence, Tuple

import numpy as np

from .checkpoint import check_model
import itertools

from .profiles import get_profile, update_config, update_profile
from .reads import get_reads, get_reads_batches
from .signal_io import BLOW5Writer, POD5Writer

logger = logging.getLogger("seq2squiggle")

BATCH_CHUNKS = int(os.environ.get("S2S_READ_BATCH_CHUNKS", 131072))  # chunks per predict_reads() call


def get_writer(out, profile, ideal_mode, export_every_n_samples, profile_name, preserve_read_ids, remove_existing=True):
    """inference.py:28-82: writer by extension; an existing output file is deleted (``remove_existing=False``: the
    caller — rank 0 of a sharded run — has done that already)."""
    out = str(out)
    out_base = os.path.basename(out)
    out_dir = os.path.dirname(out)
    if out_dir and not os.path.exists(out_dir):
        os.makedirs(out_dir, exist_ok=True)
    if remove_existing and os.path.exists(out):
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


def plan_batches(chunk_counts: Sequence[int], batch_chunks: Optional[int] = None) -> List[Tuple[int, int]]:
    """Read ranges ``[lo, hi)`` of the batches of a sharded run, in read order: a batch is closed by the read that
    brings it to ``batch_chunks`` chunks (a read is never split) — the rule of ``predict_reads``' pipeline pieces, so one
    batch is one piece.  Batch ``b`` is simulated and written by rank ``b % world``."""
    if batch_chunks is None:
        from .model import PIPE_CHUNKS
        batch_chunks = PIPE_CHUNKS
    counts = np.asarray(chunk_counts, dtype=np.int64)
    plan, lo = [], 0
    cum = np.concatenate([[0], np.cumsum(counts)])
    while lo < len(counts):
        hi = int(np.searchsorted(cum, cum[lo] + batch_chunks, side="left"))     # first hi with sum(lo..hi-1) >= batch_chunks
        hi = min(max(hi, lo + 1), len(counts))
        plan.append((lo, hi))
        lo = hi
    return plan


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
    if (world > 1 and "CUDA_VISIBLE_DEVICES" not in os.environ and not torch.cuda.is_initialized()
            and local < torch.cuda.device_count()):      # (device_count() before initialisation asks NVML, not CUDA)
        # one process per GPU: let each rank see only its own device.  CUDA initialisation enumerates (and sets up) every
        # visible device; eight ranks doing that for eight GPUs each at the same moment cost ~7 s of the 12 s whole-command
        # wall time of an 8-GPU run (profiles/r02_config5_8gpu.txt)
        os.environ["CUDA_VISIBLE_DEVICES"] = str(local)
        local = 0
    if world > 1:
        # Sharded run (one process per GPU, torchrun): the batches of the read list are dealt to the ranks round-robin and
        # every rank writes its batches straight into the ONE output file, at offsets the ranks hand each other in read
        # order (signal_io.SharedOrder) — no part files, no splice pass after the compute.
        if not (out.endswith(".blow5") or out.endswith(".slow5")):
            raise ValueError("multi-GPU predict writes one shared SLOW5/BLOW5 file: use a .blow5 / .slow5 output")
        if rank == 0 and os.path.exists(out):
            logger.warning(f"Output file {out} already exists. File will be deleted.")
            os.remove(out)
    writer, export_every_n_samples = get_writer(out, profile_dict, ideal_mode, export_every_n_samples, profile_name=profile,
                                                preserve_read_ids=preserve_read_ids, remove_existing=world == 1)
    if saved_weights is None:
        saved_weights = get_saved_weights(profile)

    local = local % max(torch.cuda.device_count(), 1)   # more ranks than GPUs (tests): ranks share devices
    plan_state = None
    if world > 1:
        # CUDA initialisation of N processes at once takes seconds (5-6 s for eight, whatever each of them may see); the
        # batch plan needs no GPU (genome preprocessing + lengths-only replay of the sampler: 1-2 s), so it runs meanwhile
        import threading
        import torch.distributed as dist
        init = threading.Thread(target=torch.cuda.set_device, args=(local,), daemon=True)
        init.start()
        if not dist.is_initialized():
            dist.init_process_group("gloo")      # control plane only: barriers around the shared file
        plan_state = get_reads_batches(fasta, read_input, n, r, c, config, distr,
                                       seed, profile, min_read_len, rank, world, plan_batches, chunks_of_read,
                                       cheap_names=not preserve_read_ids)
        init.join()
    torch.cuda.set_device(local)
    load_model = seq2squiggle.load_from_checkpoint(
        checkpoint_path=saved_weights, out_writer=writer, dwell_mean=dwell_mean, dwell_std=dwell_std,
        noise_std=noise_std, noise_sampling=noise_sampling, duration_sampling=duration_sampling,
        export_every_n_samples=export_every_n_samples, min_noise=min_noise, min_duration=min_duration, device=local,
        precision=precision)
    check_model(load_model.hparams.config, config)

    # reads are sampled lazily, batch by batch, while the GPU works on the previous batches (the sampler is
    # sequential Python); a sharded multi-process run has replayed the sampler for the read lengths alone (above), cut the
    # list into batches and materialises only its own batches (reads.get_reads_batches)
    k = config["seq_kmer"]
    if world > 1:
        import torch.distributed as dist
        from .signal_io import SharedOrder
        reads, plan, counts = plan_state
        chunk_cum = np.concatenate([[0], np.cumsum(counts)])
        mine = list(range(rank, len(plan), world))
        logger.info(f"rank {rank}/{world}: {len(mine)} of {len(plan)} batches, {len(counts)} reads in the run")
        order_path = out + ".order"
        if rank == 0:
            shared = SharedOrder(order_path, len(plan), create=True)
        dist.barrier()
        if rank != 0:
            shared = SharedOrder(order_path, len(plan), create=False)
        np.random.seed(seed % (2 ** 32))       # every rank replays the ONE per-record draw stream (offset / median_before)
        writer.begin_shared(shared, rank)
        n_reads = 0
        it = iter(reads)
        for b in mine:
            lo, hi = plan[b]
            batch = list(itertools.islice(it, hi - lo))
            load_model.predict_reads(batch, chunk_id_base=int(chunk_cum[lo]), tag=(b, lo))
            n_reads += len(batch)
        load_model.on_predict_epoch_end()
        writer.end_shared()
        stats = getattr(load_model, "last_stats", None)
        logger.info(f"rank {rank}: simulated {n_reads} reads, {writer.samples_written} samples -> {out}")
        dist.barrier()
        if rank == 0:
            os.remove(order_path)
        return stats
    reads, total_l = get_reads(fasta, read_input, n, r, c, config, distr, seed, profile, min_read_len,
                               stream=True, cheap_names=not preserve_read_ids)
    n_reads = 0
    for batch in batch_reads(reads, k):
        load_model.predict_reads(batch)
        n_reads += len(batch)
    load_model.on_predict_epoch_end()
    stats = getattr(load_model, "last_stats", None)
    logger.info(f"simulated {n_reads} reads, {writer.samples_written} samples -> {writer.filename}")
    return stats
