"""``seq2squiggle predict`` command line (reference ``seq2squiggle.py:43-84, 230-599, 640-657``).

Same argument, option names and defaults as the reference's ``predict`` command, built on plain ``click``
(``rich_click`` only styles the help text).  Both spellings found in the wild are accepted: the code's
``--noise-sampler/--duration-sampler`` and the README's ``--noise-sampling/--duration-sampling``; profile names may
use ``-`` or ``_``.  ``preprocess``, ``train`` and ``sweep`` are outside this package's scope and say so.

    python -m seq2squiggle_b200 predict example.fasta -o out.blow5 -m model.ckpt --profile dna-r10-prom -n 1000
"""
from __future__ import annotations

import logging
import os
import pathlib
import random
import sys

import click
import numpy as np

from . import __version__
from .checkpoint import set_config
from .profiles import PROFILE_NAMES, normalise_profile_name

logger = logging.getLogger("seq2squiggle")


def setup_logging(verbosity: str) -> None:
    """utils.py:687-719."""
    levels = {"debug": logging.DEBUG, "info": logging.INFO, "warning": logging.WARNING, "error": logging.ERROR}
    logging.captureWarnings(True)
    root = logging.getLogger()
    root.setLevel(logging.DEBUG)
    handler = logging.StreamHandler(sys.stderr)
    handler.setLevel(levels[verbosity.lower()])
    handler.setFormatter(logging.Formatter("{name} {levelname} {asctime}: {message}", style="{", datefmt="%H:%M:%S"))
    root.addHandler(handler)
    logging.getLogger("py.warnings").addHandler(handler)
    for name in ("fsspec", "github", "h5py", "numba", "pytorch_lightning", "torch", "urllib3"):
        logging.getLogger(name).setLevel(logging.WARNING)


def set_seeds(seed: int) -> int:
    """utils.py:722-741: seed 0 means "draw a random seed"; seeds ``random``, NumPy and torch (the device Philox
    streams are keyed by ``torch.initial_seed()``)."""
    import torch
    if not seed:
        seed = resolve_random_seed()
        logger.info(f"No seed provided. Generated random seed: {seed}")
    logger.info(f"Setting all random seeds to {seed}")
    os.environ["PYTHONHASHSEED"] = str(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    return seed


def resolve_random_seed() -> int:
    """``-s 0``: one random seed for the whole run.  Under ``torchrun`` every rank derives the read list, the shard
    bounds, the read numbers and the device Philox key from the seed, so rank 0 draws it and broadcasts it over the
    gloo control plane; ranks drawing their own would simulate different (overlapping / missing) reads."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    seed = int.from_bytes(os.urandom(4), byteorder="big", signed=False) or 1
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("gloo")
        box = [seed]
        dist.broadcast_object_list(box, src=0)
        seed = int(box[0])
    return seed


class _Profile(click.ParamType):
    name = "profile"

    def convert(self, value, param, ctx):
        v = normalise_profile_name(str(value))
        if v not in PROFILE_NAMES:
            self.fail(f"{value!r} is not one of {', '.join(PROFILE_NAMES)}", param, ctx)
        return v


@click.group(context_settings=dict(help_option_names=["-h", "--help"]))
def main():
    """seq2squiggle (B200-native predict path): nanopore signal simulation with a feed-forward transformer."""


def _shared(f):  # seq2squiggle.py:43-84
    f = click.option("-v", "--verbosity", type=click.Choice(["debug", "info", "warning", "error"], case_sensitive=False),
                     default="info", help="Verbosity of console logging messages.")(f)
    f = click.option("-y", "--config", default=None, help="The YAML configuration file overriding the default options.")(f)
    f = click.option("-m", "--model", type=click.Path(exists=False, dir_okay=False), default=None,
                     help="The model weights (.ckpt file).")(f)
    f = click.option("-s", "--seed", type=int, default=0, help="Set the seed value for reproducibility")(f)
    return f


def _advanced(f):  # seq2squiggle.py:230-390 (hidden unless --show-advanced-options)
    o = lambda *a, **kw: click.option(*a, show_default=True, hidden=True, **kw)  # noqa: E731
    f = o("--noise-sampler", "--noise-sampling", "noise_sampler", default=True, type=bool,
          help="Enable or disable the noise sampler.")(f)
    f = o("--duration-sampler", "--duration-sampling", "duration_sampler", default=True, type=bool,
          help="Enable or disable the duration sampler.")(f)
    f = o("--dwell-mean", default=None, type=float, help="Mean dwell time (signal points per k-mer); used only if the "
          "duration sampler is deactivated.")(f)
    f = o("--dwell-std", default=0.0, type=float, help="Standard deviation of the dwell time; used only if the duration "
          "sampler is deactivated.")(f)
    f = o("--noise-std", default=2.0, type=float, help="Set the standard deviation for noise.")(f)
    f = o("--distr", default="expon", type=click.Choice(["expon", "beta", "gamma"]), help="Distribution for read sampling.")(f)
    f = o("--predict-batch-size", default=1024, type=int, help="Batch size for prediction.")(f)
    f = o("--export-every-n-samples", default=1000000, type=int, help="How often the predicted samples are saved.")(f)
    f = o("--sample-rate", default=None, type=int, help="Specify the sampling rate.")(f)
    f = o("--bps", default=None, type=int, help="Specify the translocation speed.")(f)
    f = o("--digitisation", default=None, type=int, help="Specify the digitisation.")(f)
    f = o("--range_val", default=None, type=float, help="Specify the range value.")(f)
    f = o("--offset_mean", default=None, type=float, help="Specify the offset mean.")(f)
    f = o("--offset_std", default=None, type=float, help="Specify the offset standard deviation.")(f)
    f = o("--median_before_mean", default=None, type=float, help="Specify the median_before mean.")(f)
    f = o("--median_before_std", default=None, type=float, help="Specify the median_before standard deviation.")(f)
    f = o("--min_noise", default=0.0, type=float, help="Minimal stdv value for the noise sampler.")(f)
    f = o("--min_duration", default=3, type=int, help="Minimal event duration.")(f)
    f = o("--min_read_len", default=30, type=int, help="Minimal read length for reference mode.")(f)
    f = o("--precision", default="fp16", type=click.Choice(["fp16", "fp32"]),
          help="B200 engine arithmetic: fp16 tensor-core path (the reference's own 16-mixed GPU mode) or fp32 parity path.")(f)
    f = click.option("--preserve-read-ids", is_flag=True, default=False, show_default=True,
                     help="Preserve original read IDs from input instead of generating synthetic UUID4s.")(f)
    return f


@main.command(context_settings={"ignore_unknown_options": True})
@click.argument("fasta", required=False, type=click.Path(exists=False, file_okay=True, dir_okay=False, path_type=pathlib.Path))
@click.option("--read-input", default=False, is_flag=True, show_default=True,
              help="Enable Read Mode: simulate signals directly from input reads in a FASTA or FASTQ file.")
@click.option("-n", "--num-reads", type=int, default=-1, help="Specify the desired number of generated reads.")
@click.option("-r", "--read-length", type=int, default=1000, show_default=True, help="Specify the desired average read length.")
@click.option("-c", "--coverage", type=int, default=-1, help="Specify the desired genome coverage.")
@click.option("-o", "--out", required=False, type=click.Path(file_okay=True, dir_okay=False, path_type=pathlib.Path),
              help="Specify the path to the output POD5/SLOW5/BLOW5 file.")
@click.option("--profile", default="dna-r10-prom", show_default=True, type=_Profile(),
              help="Select a profile for data simulation: " + ", ".join(PROFILE_NAMES))
@click.option("--show-advanced-options", is_flag=True, default=False, help="Show advanced options for signal prediction.")
@_advanced
@_shared
@click.pass_context
def predict(ctx, fasta, read_input, num_reads, read_length, coverage, out, profile, show_advanced_options, noise_sampler,
            duration_sampler, dwell_mean, dwell_std, noise_std, distr, predict_batch_size, export_every_n_samples,
            sample_rate, bps, digitisation, range_val, offset_mean, offset_std, median_before_mean, median_before_std,
            min_noise, min_duration, min_read_len, precision, preserve_read_ids, seed, model, config, verbosity):
    """Generate sequencing signals from genome or read fasta file

    FASTA must be .fasta file with desired genome or reads for simulation
    """
    if show_advanced_options:
        for param in ctx.command.params:
            param.hidden = False
        click.echo(ctx.get_help())
        ctx.exit()
    if not fasta or not out:
        logger.error("FASTA file and Output file are required for prediction.")
        ctx.exit(1)
    setup_logging(verbosity)
    logger.info("seq2squiggle_b200 version %s", str(__version__))
    args = dict(fasta=fasta, read_input=read_input, num_reads=num_reads, read_length=read_length, coverage=coverage,
                out=out, profile=profile, noise_sampler=noise_sampler, duration_sampler=duration_sampler,
                dwell_mean=dwell_mean, dwell_std=dwell_std, noise_std=noise_std, distr=distr,
                predict_batch_size=predict_batch_size, export_every_n_samples=export_every_n_samples,
                sample_rate=sample_rate, bps=bps, digitisation=digitisation, range=range_val, offset_mean=offset_mean,
                offset_std=offset_std, median_before_mean=median_before_mean, median_before_std=median_before_std,
                min_noise=min_noise, min_duration=min_duration, min_read_len=min_read_len,
                preserve_read_ids=preserve_read_ids, seed=seed, model=model, config=config, verbosity=verbosity,
                precision=precision)
    logger.info("Arguments:")
    for key, value in args.items():
        logger.info(f" {key}: {value}")
    config = set_config(config)
    logger.debug("Config parameters:")
    for key in config:
        logger.debug(f" {key}: {config[key]}")
    seed = set_seeds(seed)

    from .inference import inference_run
    inference_run(config=config, saved_weights=model, fasta=fasta, read_input=read_input, n=num_reads, r=read_length,
                  c=coverage, out=out, profile=profile, dwell_mean=dwell_mean, dwell_std=dwell_std, noise_std=noise_std,
                  noise_sampling=noise_sampler, duration_sampling=duration_sampler, distr=distr,
                  predict_batch_size=predict_batch_size, export_every_n_samples=export_every_n_samples,
                  sample_rate=sample_rate, bps=bps, digitisation=digitisation, range_val=range_val,
                  offset_mean=offset_mean, offset_std=offset_std, median_before_mean=median_before_mean,
                  median_before_std=median_before_std, min_noise=min_noise, min_duration=min_duration,
                  min_read_len=min_read_len, preserve_read_ids=preserve_read_ids, seed=seed, precision=precision)
    logger.info("Prediction done.")


def _out_of_scope(name):
    @main.command(name=name, context_settings={"ignore_unknown_options": True, "allow_extra_args": True})
    def cmd():
        raise click.ClickException(f"'{name}' is not part of seq2squiggle_b200 (predict path only); use the reference "
                                   "seq2squiggle package for it.")
    cmd.__doc__ = f"Not available here: '{name}' belongs to the reference package."
    return cmd


for _name in ("preprocess", "train", "sweep"):
    _out_of_scope(_name)


@main.command()
def version():
    """Get the version of seq2squiggle_b200"""
    import torch
    setup_logging("info")
    logger.info(f"seq2squiggle_b200: {__version__}")
    logger.info(f"pytorch: {torch.__version__}")


if __name__ == "__main__":
    main()
