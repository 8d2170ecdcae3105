"""CPU restatement of the ``seq2squiggle predict`` signal-generation path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function cites the
reference lines it follows (paths relative to ``/root/reference/src/seq2squiggle``).
The arithmetic is stated with ``torch`` CPU functional ops in the same order the
reference modules issue them, on a plain ``state_dict`` (no ``nn.Module``), so
that in fp32 / matmul-precision "highest" it reproduces the reference bit for
bit; ``tests/test_oracle_golden.py`` pins that against vectors produced by the
reference's own modules (``oracle/make_golden.py``).
"""
from __future__ import annotations

from collections import OrderedDict, defaultdict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- #
# constants (config.yaml:14-33)
# --------------------------------------------------------------------------- #
DEFAULT_CONFIG = {
    "scaling_max_value": 165.0,
    "max_dna_len": 16,
    "max_signal_len": 250,
    "allowed_chars": "_ACGT",
    "seq_kmer": 9,
    "pre_layers": 1,
    "dmodel": 64,
    "dff": 256,
    "encoder_layers": 2,
    "encoder_heads": 8,
    "decoder_layers": 2,
    "decoder_heads": 8,
    "encoder_dropout": 0.2,
    "decoder_dropout": 0.2,
    "duration_dropout": 0.2,
}

LETTER_TO_INT = {"_": 0, "A": 1, "C": 2, "G": 3, "T": 4}  # utils.py:74


# --------------------------------------------------------------------------- #
# a1 tokeniser  (utils.py:56-89, 266-287, 334-356)
# --------------------------------------------------------------------------- #
def extract_kmers(dna_string: str, k: int) -> List[str]:
    """utils.py:334-339 — every overlapping k-mer (n-k+1 of them)."""
    return [dna_string[i:i + k] for i in range(len(dna_string) - k + 1)]


def add_remainder(x: List[str], max_dna: int, k: int) -> List[str]:
    """utils.py:342-347 — right-pad the k-mer list to a multiple of max_dna with '_'*k."""
    remain = max_dna - (len(x) % max_dna)
    if remain % max_dna > 0:
        x = x + [("_" * k)] * remain
    return x


def one_hot_encode(sequences: Sequence[str], seq_len: int) -> np.ndarray:
    """utils.py:56-89 — float16 [n, k, 5]; letters outside '_ACGT' give an all-zero row."""
    out = np.zeros((len(sequences), seq_len, 5), dtype=np.float16)
    for i, kmer in enumerate(sequences):
        for j, letter in enumerate(kmer):
            idx = LETTER_TO_INT.get(letter)
            if idx is not None:
                out[i, j, idx] = 1
    return out


def regular_break_points(n: int, chunk_len: int, overlap: int = 0, align: str = "left") -> np.ndarray:
    """utils.py:266-287."""
    num_chunks, remainder = divmod(n - overlap, chunk_len - overlap)
    start = {"left": 0, "mid": remainder // 2, "right": remainder}[align]
    starts = np.arange(start, start + num_chunks * (chunk_len - overlap), (chunk_len - overlap))
    return np.vstack([starts, starts + chunk_len]).T


def split_sequence(x: str, config: dict) -> np.ndarray:
    """utils.py:350-356 — read string -> float16 [n_chunks, 16, k, 5]."""
    k = config["seq_kmer"]
    kmers = extract_kmers(x, k)
    kmers = add_remainder(kmers, config["max_dna_len"], k)
    oh = one_hot_encode(kmers, k)
    bps = regular_break_points(len(oh), config["max_dna_len"], align="left")
    if len(bps) == 0:
        return np.zeros((0, config["max_dna_len"], k, 5), dtype=np.float16)
    return np.array([oh[i:j] for (i, j) in bps])


def split_sequence_fast(x: str, config: dict) -> np.ndarray:
    """Vectorised equivalent of :func:`split_sequence` (same output, used to build
    large oracle inputs in seconds; checked against the loop version in tests)."""
    k = config["seq_kmer"]
    L = config["max_dna_len"]
    n = len(x) - k + 1
    if n <= 0:
        return np.zeros((0, L, k, 5), dtype=np.float16)
    lut = np.full(256, -1, dtype=np.int8)
    for ch, i in LETTER_TO_INT.items():
        lut[ord(ch)] = i
    codes = lut[np.frombuffer(x.encode("latin-1", "replace"), dtype=np.uint8)]
    n_pad = (-n) % L
    idx = np.arange(n)[:, None] + np.arange(k)[None, :]
    kc = codes[idx]  # [n, k]
    if n_pad:
        kc = np.concatenate([kc, np.zeros((n_pad, k), dtype=np.int8)], 0)
    oh = np.zeros((kc.shape[0], k, 5), dtype=np.float16)
    valid = kc >= 0
    ii, jj = np.nonzero(valid)
    oh[ii, jj, kc[ii, jj]] = 1
    return oh.reshape(-1, L, k, 5)


def n_chunks_of_read(read_len: int, k: int, max_dna: int = 16) -> int:
    n = read_len - k + 1
    return 0 if n <= 0 else -(-n // max_dna)


# --------------------------------------------------------------------------- #
# a4 FFT block (layers.py:11-41, 44-88, 91-113, 116-142)
# --------------------------------------------------------------------------- #
def _mha(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor, n_head: int) -> torch.Tensor:
    """layers.py:64-88 with mask=None (model.py:217) and dropout inert (eval)."""
    sz_b, L, d_model = x.shape
    d_k = d_model // n_head
    residual = x
    q = F.linear(x, sd[p + "w_qs.weight"], sd[p + "w_qs.bias"]).view(sz_b, L, n_head, d_k)
    k = F.linear(x, sd[p + "w_ks.weight"], sd[p + "w_ks.bias"]).view(sz_b, L, n_head, d_k)
    v = F.linear(x, sd[p + "w_vs.weight"], sd[p + "w_vs.bias"]).view(sz_b, L, n_head, d_k)
    q = q.permute(2, 0, 1, 3).contiguous().view(-1, L, d_k)
    k = k.permute(2, 0, 1, 3).contiguous().view(-1, L, d_k)
    v = v.permute(2, 0, 1, 3).contiguous().view(-1, L, d_k)
    attn = torch.bmm(q, k.transpose(1, 2))          # layers.py:20
    attn = attn / (d_k ** 0.5)                      # layers.py:21 (temperature = d_k**0.5, :58)
    attn = torch.softmax(attn, dim=2)               # layers.py:39
    out = torch.bmm(attn, v)                        # layers.py:40
    out = out.view(n_head, sz_b, L, d_k).permute(1, 2, 0, 3).contiguous().view(sz_b, L, -1)
    out = F.linear(out, sd[p + "fc.weight"], sd[p + "fc.bias"])
    out = F.layer_norm(out + residual, (d_model,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)
    return out


def _ffn(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor) -> torch.Tensor:
    """layers.py:108-113."""
    residual = x
    h = F.relu(F.linear(x, sd[p + "w_1.weight"], sd[p + "w_1.bias"]))
    out = F.linear(h, sd[p + "w_2.weight"], sd[p + "w_2.bias"])
    return F.layer_norm(out + residual, (x.shape[-1],), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)


def fft_block(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor, n_head: int) -> torch.Tensor:
    """layers.py:135-142."""
    return _ffn(sd, p + "pos_ffn.", _mha(sd, p + "slf_attn.", x, n_head))


# --------------------------------------------------------------------------- #
# a3 encoder, a9 decoder, a5/a6 samplers (modules.py)
# --------------------------------------------------------------------------- #
def encoder_forward(sd, config, src: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """modules.py:65-89.  src: [B,16,5k] (any float dtype) -> (enc_out, emb_out)."""
    x = src.float()
    x = F.relu(F.linear(x, sd["encoders.src_emb.weight"], sd["encoders.src_emb.bias"]))
    for i in range(config["pre_layers"]):
        x = F.relu(F.linear(x, sd[f"encoders.pre_net_stack.{i}.weight"], sd[f"encoders.pre_net_stack.{i}.bias"]))
    emb_out = x
    # position_enc is [1,16,64]; the [:L] slice on dim 0 is a no-op (modules.py:80)
    enc = x + sd["encoders.position_enc"][: x.shape[1]]
    for i in range(config["encoder_layers"]):
        enc = fft_block(sd, f"encoders.layer_stack.{i}.", enc, config["encoder_heads"])
    return enc, emb_out


def decoder_forward(sd, config, x: torch.Tensor) -> torch.Tensor:
    """modules.py:133-142.  x: [B,250,64] -> [B,250,1] >= 0."""
    y = x + sd["decoders.position_enc"][: x.shape[1]]
    for i in range(config["decoder_layers"]):
        y = fft_block(sd, f"decoders.layer_stack_FFT.{i}.", y, config["decoder_heads"])
    y = F.linear(y, sd["decoders.out_linear.weight"], sd["decoders.out_linear.bias"])
    return F.relu(y)


def _softplus_mlp(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """Linear-ReLU-(Dropout)-Linear-Softplus (modules.py:180-193, 266-272)."""
    h = F.relu(F.linear(x, sd[p + "0.weight"], sd[p + "0.bias"]))
    return F.softplus(F.linear(h, sd[p + "3.weight"], sd[p + "3.bias"]))


def noise_sampler_forward(sd, emb_out: torch.Tensor) -> torch.Tensor:
    """modules.py:275-278 -> [B,16]."""
    return _softplus_mlp(sd, "noise_sampler.stdv_layer.", emb_out).flatten(1)


def duration_params(sd, emb_out: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """modules.py:216-219 — (conc, rate), each clamped to >= 1e-8, shape [B,16,1]."""
    conc = torch.clamp(_softplus_mlp(sd, "length_regulator.duration_sampler.conc_layer.", emb_out), min=1e-8)
    rate = torch.clamp(_softplus_mlp(sd, "length_regulator.duration_sampler.rate_layer.", emb_out), min=1e-8)
    return conc, rate


def duration_sampler_forward(sd, emb_out: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """modules.py:197-225 — Gamma(conc, rate).sample(), clamp >= 1.0, flatten."""
    conc, rate = duration_params(sd, emb_out)
    if generator is None:
        out = torch.distributions.gamma.Gamma(concentration=conc, rate=rate).sample()
    else:  # same law, explicit generator (torch._standard_gamma(conc)/rate, gamma.py rsample)
        out = torch._standard_gamma(conc, generator=generator) / rate
        out = out.clamp(min=torch.finfo(out.dtype).tiny)
    return torch.clamp(out, min=1.0).flatten(1)


# --------------------------------------------------------------------------- #
# a7/a8 length regulator (modules.py:344-441)
# --------------------------------------------------------------------------- #
def lr_expand(x: torch.Tensor, x_noise: Optional[torch.Tensor], dur: torch.Tensor, max_length: int):
    """modules.py:344-392, literal restatement (alignment matrix + bmm + F.pad)."""
    bsz, n_in = dur.shape
    cum = torch.cumsum(dur, dim=1)
    t_max = int(torch.max(cum))
    ids = torch.arange(t_max)
    m = (ids.unsqueeze(0) < cum.reshape(bsz * n_in).unsqueeze(1)).reshape(bsz, n_in, t_max).float()
    m = torch.diff(m, dim=1, prepend=torch.zeros_like(m[:, :1]))
    out = torch.bmm(m.permute(0, 2, 1), x)
    if x_noise is not None:
        x_noise = torch.bmm(m.permute(0, 2, 1), x_noise)
    if max_length:
        out = F.pad(out, (0, 0, 0, max_length - out.size(1), 0, 0))
        if x_noise is not None:
            x_noise = F.pad(x_noise, (0, 0, 0, max_length - x_noise.size(1), 0, 0))
    return out, x_noise


def lr_expand_indices(dur: np.ndarray, max_length: int) -> np.ndarray:
    """Integer form of modules.py:344-392: src k-mer index per output position, -1 = zero fill.
    Equivalent to repeat_interleave(arange(16), dur)[:max_length]."""
    dur = np.asarray(dur, dtype=np.int64)
    bsz, n_in = dur.shape
    cum = np.cumsum(dur, axis=1)
    t = np.arange(max_length)[None, :, None]
    # j(t) = number of cum values <= t
    j = (cum[:, None, :] <= t).sum(-1)
    j[j >= n_in] = -1
    return j.astype(np.int32)


def durations_forward(sd, emb_out, *, dwell_mean, dwell_std, duration_sampling, min_length,
                      generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """modules.py:410-437 — returns (float durations [B,16], int32 rounded durations)."""
    if duration_sampling:
        d = duration_sampler_forward(sd, emb_out.detach().clone(), generator)
        d = torch.clamp(d, min=min_length)
    else:
        bs, seq, _ = emb_out.shape
        if dwell_std <= 0:
            d = torch.full((bs, seq), dwell_mean)            # note: no min_length clamp here
        else:
            mean = torch.full((bs, seq), dwell_mean)
            std = torch.full((bs, seq), dwell_std)
            d = torch.normal(mean=mean, std=std, generator=generator)
            d = torch.clamp(d, min=min_length)
    return d, torch.round(d.detach().clone()).int()           # half-to-even


# --------------------------------------------------------------------------- #
# a2 + a10 predict step (model.py:195-240)
# --------------------------------------------------------------------------- #
def predict_step(sd, config, data: torch.Tensor, *, dwell_mean: float, dwell_std: float = 0.0,
                 noise_std: float = 0.0, noise_sampling: bool = False, duration_sampling: bool = False,
                 min_noise: float = 0.0, min_duration: float = 1, generator: Optional[torch.Generator] = None,
                 return_stages: bool = False):
    """model.py:195-240.  data: [B,16,k,5] one-hot.  Returns pA [B,250] (and the stage tensors)."""
    bs, seq_l = data.shape[:2]
    data = data.reshape(bs, seq_l, -1)
    enc_out, emb_out = encoder_forward(sd, config, data)
    sigma = noise_sampler_forward(sd, emb_out)[:, :, None]
    dur_f, dur_i = durations_forward(sd, emb_out, dwell_mean=dwell_mean, dwell_std=dwell_std,
                                     duration_sampling=duration_sampling, min_length=min_duration,
                                     generator=generator)
    lr_out, sigma_ext = lr_expand(enc_out, sigma, dur_i, config["max_signal_len"])
    p = decoder_forward(sd, config, lr_out)
    pred = (p * config["scaling_max_value"]).squeeze(-1)
    if noise_std > 0:
        nz = pred != 0
        if noise_sampling:
            s = torch.clamp(sigma_ext, min=min_noise).squeeze(-1) * noise_std * config["scaling_max_value"]
            g = torch.normal(mean=torch.zeros_like(s), std=s, generator=generator)
            pred[nz] += g[nz]
        else:
            g = torch.normal(mean=0.0, std=float(noise_std), size=pred.shape, generator=generator)
            pred[nz] += g[nz]
    pred = torch.clamp(pred, min=0)
    if return_stages:
        return pred, dict(emb_out=emb_out, enc_out=enc_out, sigma=sigma.squeeze(-1), dur_f=dur_f, dur_i=dur_i,
                          lr_out=lr_out, sigma_ext=sigma_ext.squeeze(-1), p=p.squeeze(-1))
    return pred


# --------------------------------------------------------------------------- #
# a11 result assembly (model.py:242-307) and a12 digitisation (signal_io.py:134-141)
# --------------------------------------------------------------------------- #
def assemble_reads(read_ids: Sequence[str], pred_rows: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
    """model.py:242-245 + 262-286 for one flush with keep_last=False: group rows by read id in
    first-seen order, concatenate, drop every exact zero."""
    res: Dict[str, list] = defaultdict(list)
    for rid, row in zip(read_ids, pred_rows):
        res[rid].append(row)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in res.items():
        cat = torch.cat(v)
        out[k] = cat[cat.nonzero()].squeeze()
    return out


def digitise(signal_pa: np.ndarray, digitisation: float, signal_range: float, offset: float, rna: bool = False) -> np.ndarray:
    """signal_io.py:134-141: float32 signal, python-float scalars (NumPy keeps float32),
    multiply, divide, subtract, np.round (half-even), astype(int16) (wraps), RNA reversed."""
    signal = np.asarray(signal_pa).astype(np.float32)
    raw = np.round(signal * np.float32(digitisation) / np.float32(signal_range) - np.float32(offset))
    with np.errstate(invalid="ignore"):
        raw = raw.astype(np.int64).astype(np.int16)  # two's-complement wrap like the C cast on x86
    if rna:
        raw = np.ascontiguousarray(raw[::-1])
    return raw


# --------------------------------------------------------------------------- #
# checkpoint: random-init default architecture in the reference's layout
# (model.py:46-50 construction order; nn.Linear / nn.LayerNorm default init)
# --------------------------------------------------------------------------- #
def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """layers.py:145-165: angles in Python float64, tensor in float32, sin/cos in float32."""
    tab = torch.tensor([[pos / 10000 ** (2 * (j // 2) / d_hid) for j in range(d_hid)] for pos in range(n_position)])
    tab[:, 0::2] = torch.sin(tab[:, 0::2])
    tab[:, 1::2] = torch.cos(tab[:, 1::2])
    return torch.FloatTensor(tab)


def random_init_state_dict(config: dict, seed: int) -> "OrderedDict[str, torch.Tensor]":
    """State dict with the key names/shapes of ``seq2squiggle(config).state_dict()``; parameters are
    drawn by constructing torch.nn layers under ``torch.manual_seed(seed)`` in the reference's
    construction order (model.py:47-50 -> modules.py:37-62, 112-132, 170-193, 266-272;
    layers.py:54-62, 96-106) so the values equal a reference model built under the same seed."""
    import torch.nn as nn
    d, dff = config["dmodel"], config["dff"]
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def lin(name, fin, fout):
        m = nn.Linear(fin, fout)
        sd[name + ".weight"] = m.weight.detach().clone()
        sd[name + ".bias"] = m.bias.detach().clone()

    def ln(name):
        sd[name + ".weight"] = torch.ones(d)
        sd[name + ".bias"] = torch.zeros(d)

    def fft(prefix):
        for w in ("w_qs", "w_ks", "w_vs"):
            lin(prefix + "slf_attn." + w, d, d)
        ln(prefix + "slf_attn.layer_norm")
        lin(prefix + "slf_attn.fc", d, d)
        lin(prefix + "pos_ffn.w_1", d, dff)
        lin(prefix + "pos_ffn.w_2", dff, d)
        ln(prefix + "pos_ffn.layer_norm")

    g = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        sd["encoders.position_enc"] = sinusoid_table(config["max_dna_len"], d).unsqueeze(0)
        lin("encoders.src_emb", len(config["allowed_chars"]) * config["seq_kmer"], d)
        for i in range(config["pre_layers"]):
            lin(f"encoders.pre_net_stack.{i}", d, d)
        for i in range(config["encoder_layers"]):
            fft(f"encoders.layer_stack.{i}.")
        for name in ("conc_layer", "rate_layer"):
            lin(f"length_regulator.duration_sampler.{name}.0", d, d)
            lin(f"length_regulator.duration_sampler.{name}.3", d, 1)
        sd["decoders.position_enc"] = sinusoid_table(config["max_signal_len"], d).unsqueeze(0)
        lin("decoders.out_linear", d, 1)
        for i in range(config["decoder_layers"]):
            fft(f"decoders.layer_stack_FFT.{i}.")
        lin("noise_sampler.stdv_layer.0", d, d)
        lin("noise_sampler.stdv_layer.3", d, 1)
    finally:
        torch.random.set_rng_state(g)
    return sd


def lightning_checkpoint(sd, config: dict, **hparams) -> dict:
    """The dict ``Trainer.save_checkpoint(weights_only=True)`` writes for the reference model
    (SURVEY §5 'Checkpoint'): state_dict + hyper_parameters (model.py:30-46 keyword args)."""
    hp = dict(config=dict(config), save_valid_plots=True, out_writer=None, dwell_mean=9.0, dwell_std=0.0,
              noise_std=-1, noise_sampling=False, duration_sampling=False, export_every_n_samples=2000000,
              min_noise=0.5, min_duration=1)
    hp.update(hparams)
    return {
        "epoch": 0,
        "global_step": 0,
        "pytorch-lightning_version": "2.5.1.post0",
        "state_dict": OrderedDict((k, v.clone()) for k, v in sd.items()),
        "loops": {},
        "hparams_name": "kwargs",
        "hyper_parameters": hp,
    }
