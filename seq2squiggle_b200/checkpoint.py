"""Checkpoint layout of the reference model and the packed weight blob of the C-ABI.

``seq2squiggle.load_from_checkpoint`` (inference.py:386-397) reads a ``torch.save``d Lightning dict:
``state_dict`` (84 tensors, names from model.py:47-50 / modules.py / layers.py) and ``hyper_parameters``
(``config`` = the training YAML + the keyword arguments of model.py:30-44).  No Lightning import is needed to
read it.  ``pack_weights`` flattens the state dict into the fp32 blob order of ``csrc/s2s_weights.h``.
"""
from __future__ import annotations

import logging
import os
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np
import torch
import yaml

logger = logging.getLogger("seq2squiggle")

ARCH_FIXED = dict(dmodel=64, dff=256, encoder_heads=8, decoder_heads=8, max_dna_len=16, max_signal_len=250,
                  pre_layers=1)


# Defaults of ``seq2squiggle predict`` when no ``-y`` YAML is given (values of the reference's packaged config.yaml:14-47;
# a reference YAML passed with ``-y`` and a reference checkpoint's ``hyper_parameters["config"]`` carry the same keys,
# which is what ``check_model`` walks).  Grouped by who reads them here.
DEFAULT_CONFIG = {
    # the geometry of one chunk and the architecture the kernels are compiled for (ARCH_FIXED) / sized by
    "seq_kmer": 9, "allowed_chars": "_ACGT", "max_dna_len": 16, "max_signal_len": 250, "scaling_max_value": 165.0,
    "pre_layers": 1, "dmodel": 64, "dff": 256,
    "encoder_layers": 2, "encoder_heads": 8, "decoder_layers": 2, "decoder_heads": 8,
    # inert at predict time (eval mode) but compared by check_model
    "encoder_dropout": 0.2, "decoder_dropout": 0.2, "duration_dropout": 0.2,
    # training / preprocessing / logging keys: never read on the predict path, carried so that a checkpoint written from
    # this config and the reference's own compare key by key (mismatches only log a warning, inference.py:224-267)
    "log_name": "Human-R1041-4khz", "wandb_logger_state": "disabled",
    "max_chunks_train": 210000000, "max_chunks_valid": 100000, "train_valid_split": 0.9,
    "train_batch_size": 512, "max_epochs": 25, "save_model": True, "optimizer": "Adam", "warmup_ratio": 0.01,
    "lr": 0.0005, "weight_decay": 0.0, "lr_schedule": "warmup_cosine", "gradient_clip_val": 1.0,
}


def set_config(config_path=None) -> dict:
    """seq2squiggle.py:640-657: a YAML given with ``-y``, else the packaged defaults."""
    if config_path is None:
        logger.info("Config file was not specified. Default config will be used.")
        return dict(DEFAULT_CONFIG)
    try:
        with open(config_path, "r") as fh:
            config = yaml.safe_load(fh)
    except FileNotFoundError:
        logger.error(f"Configuration file not found: {config_path}")
        raise
    except yaml.YAMLError as exc:
        logger.error(f"Error parsing YAML file: {config_path} - {exc}")
        raise
    return config


def load_checkpoint(path: str) -> Tuple["OrderedDict[str, torch.Tensor]", dict]:
    """Returns (state_dict, hyper_parameters) of a reference ``.ckpt``."""
    if not os.path.exists(path):
        raise FileNotFoundError(f"Model checkpoint not found: {path}")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    if "state_dict" not in ck:
        raise ValueError(f"{path} is not a seq2squiggle checkpoint (no 'state_dict')")
    hp = dict(ck.get("hyper_parameters", {}))
    if "config" not in hp:
        raise ValueError(f"{path} has no hyper_parameters['config']")
    return ck["state_dict"], hp


def _block_keys(prefix: str):
    a, f = prefix + "slf_attn.", prefix + "pos_ffn."
    return [a + "w_qs.weight", a + "w_qs.bias", a + "w_ks.weight", a + "w_ks.bias", a + "w_vs.weight", a + "w_vs.bias",
            a + "layer_norm.weight", a + "layer_norm.bias", a + "fc.weight", a + "fc.bias",
            f + "w_1.weight", f + "w_1.bias", f + "w_2.weight", f + "w_2.bias", f + "layer_norm.weight",
            f + "layer_norm.bias"]


def _mlp_keys(prefix: str):
    return [prefix + "0.weight", prefix + "0.bias", prefix + "3.weight", prefix + "3.bias"]


def weight_key_order(config: dict):
    """Blob order == csrc/s2s_weights.h:map_weights."""
    keys = ["encoders.position_enc", "encoders.src_emb.weight", "encoders.src_emb.bias",
            "encoders.pre_net_stack.0.weight", "encoders.pre_net_stack.0.bias"]
    for i in range(config["encoder_layers"]):
        keys += _block_keys(f"encoders.layer_stack.{i}.")
    keys += _mlp_keys("length_regulator.duration_sampler.conc_layer.")
    keys += _mlp_keys("length_regulator.duration_sampler.rate_layer.")
    keys += ["decoders.position_enc", "decoders.out_linear.weight", "decoders.out_linear.bias"]
    for i in range(config["decoder_layers"]):
        keys += _block_keys(f"decoders.layer_stack_FFT.{i}.")
    keys += _mlp_keys("noise_sampler.stdv_layer.")
    return keys


def check_architecture(config: dict) -> None:
    for k, v in ARCH_FIXED.items():
        if config.get(k) != v:
            raise ValueError(f"Unsupported architecture: {k}={config.get(k)} (the sm_100a kernels are compiled for {v})")
    if config.get("allowed_chars") != "_ACGT":
        raise ValueError(f"Unsupported allowed_chars {config.get('allowed_chars')!r}")


def pack_weights(state_dict: Dict[str, torch.Tensor], config: dict) -> np.ndarray:
    check_architecture(config)
    k = int(config["seq_kmer"])
    expect = {"encoders.position_enc": (1, 16, 64), "encoders.src_emb.weight": (64, 5 * k),
              "decoders.position_enc": (1, 250, 64), "decoders.out_linear.weight": (1, 64)}
    parts = []
    for key in weight_key_order(config):
        if key not in state_dict:
            raise KeyError(f"checkpoint state_dict is missing {key}")
        t = state_dict[key].detach().to(torch.float32).cpu()
        if key in expect and tuple(t.shape) != expect[key]:
            raise ValueError(f"{key} has shape {tuple(t.shape)}, expected {expect[key]}")
        parts.append(t.contiguous().reshape(-1).numpy())
    return np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)


def check_model(model_config: dict, config: dict) -> None:
    """inference.py:224-267: warn on mismatching parameters, raise on a seq_kmer mismatch."""
    exclude = ["log_name", "wandb_logger_state", "max_chunks_train", "max_chunks_valid", "train_valid_split",
               "train_batch_size", "save_model"]
    for param, value in config.items():
        if param in exclude:
            continue
        if model_config.get(param) != value:
            if param == "seq_kmer":
                raise ValueError(
                    f"Parameter 'seq_kmer' mismatch: Model checkpoint value is "
                    f"{model_config.get(param)}, while config value is {value}. "
                    f"The model was trained on {model_config.get(param)}-mers, while the config file expects {value}-mers. "
                    "Choose a different model or change the config value or the --profile option. ")
            logger.warning(f"Mismatching {param} parameter in model checkpoint "
                           f"({model_config.get(param)}) and in config file ({value})")


def _sinusoid(n_position: int, d: int) -> torch.Tensor:
    """layers.py:145-165."""
    tab = torch.tensor([[pos / 10000 ** (2 * (j // 2) / d) for j in range(d)] for pos in range(n_position)])
    tab[:, 0::2] = torch.sin(tab[:, 0::2])
    tab[:, 1::2] = torch.cos(tab[:, 1::2])
    return tab.float()


def random_init_checkpoint(config: dict, seed: int = 1) -> dict:
    """A random-init checkpoint of the default architecture in the reference's Lightning layout
    (used by the benchmark and the smoke test: there is no network to fetch trained weights)."""
    g = torch.Generator().manual_seed(seed)
    d, dff, k = config["dmodel"], config["dff"], config["seq_kmer"]
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def lin(name, fin, fout):  # nn.Linear default init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both
        b = 1.0 / fin ** 0.5
        sd[name + ".weight"] = (torch.rand(fout, fin, generator=g) * 2 - 1) * b
        sd[name + ".bias"] = (torch.rand(fout, generator=g) * 2 - 1) * b

    def ln(name):
        sd[name + ".weight"], sd[name + ".bias"] = torch.ones(d), torch.zeros(d)

    def fft(p):
        for w in ("w_qs", "w_ks", "w_vs"):
            lin(p + "slf_attn." + w, d, d)
        ln(p + "slf_attn.layer_norm")
        lin(p + "slf_attn.fc", d, d)
        lin(p + "pos_ffn.w_1", d, dff)
        lin(p + "pos_ffn.w_2", dff, d)
        ln(p + "pos_ffn.layer_norm")

    sd["encoders.position_enc"] = _sinusoid(config["max_dna_len"], d).unsqueeze(0)
    lin("encoders.src_emb", 5 * k, d)
    lin("encoders.pre_net_stack.0", d, d)
    for i in range(config["encoder_layers"]):
        fft(f"encoders.layer_stack.{i}.")
    for name in ("conc_layer", "rate_layer"):
        lin(f"length_regulator.duration_sampler.{name}.0", d, d)
        lin(f"length_regulator.duration_sampler.{name}.3", d, 1)
    sd["decoders.position_enc"] = _sinusoid(config["max_signal_len"], d).unsqueeze(0)
    lin("decoders.out_linear", d, 1)
    for i in range(config["decoder_layers"]):
        fft(f"decoders.layer_stack_FFT.{i}.")
    lin("noise_sampler.stdv_layer.0", d, d)
    lin("noise_sampler.stdv_layer.3", d, 1)
    hp = dict(config=dict(config), save_valid_plots=True, out_writer=None, dwell_mean=9.0, dwell_std=0.0,
              noise_std=-1, noise_sampling=False, duration_sampling=False, export_every_n_samples=2000000,
              min_noise=0.5, min_duration=1)
    return {"epoch": 0, "global_step": 0, "pytorch-lightning_version": "2.5.1.post0", "state_dict": sd,
            "loops": {}, "hparams_name": "kwargs", "hyper_parameters": hp}
