#!/usr/bin/env python
"""Generate ``tests/golden/*`` by running the UNMODIFIED reference modules.

TEST INFRASTRUCTURE.  Run in the build container only (``/root/reference`` is not
present on the GPU box): ``python oracle/make_golden.py``.  It imports
``seq2squiggle.layers`` / ``seq2squiggle.modules`` (the arithmetic of the hot path)
and ``seq2squiggle.utils`` (tokeniser, profiles, read sampler; its plotting /
pysam / prettytable imports are stubbed because those packages are absent) straight
from the read-only reference tree, feeds them seeded inputs and stores inputs +
outputs.  The ~30 glue lines of ``model.py:195-240`` / ``signal_io.py:134-141``
cannot be imported (pytorch_lightning / pyslow5 / pod5 are not installed), so they
are transcribed here around the reference's own module calls.
"""
from __future__ import annotations

import hashlib
import json
import os
import random
import sys
import types

import numpy as np
import torch
import yaml

REF_SRC = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "prettytable", "pysam"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].ticker = sys.modules["matplotlib.ticker"]
    sys.modules["matplotlib.ticker"].AutoMinorLocator = object
    sys.modules["prettytable"].PrettyTable = object
    sys.modules["pysam"].FastxFile = object
    sys.path.insert(0, REF_SRC)
    from seq2squiggle import modules as ref_modules, utils as ref_utils  # noqa
    return ref_modules, ref_utils


def build_reference_model(ref_modules, config, seed):
    """model.py:47-50 construction order under a fixed seed."""
    torch.manual_seed(seed)
    enc = ref_modules.Encoder(config)
    lr = ref_modules.LengthRegulator(config)
    dec = ref_modules.Decoder(config)
    ns = ref_modules.NoiseSampler(config)
    for m in (enc, lr, dec, ns):
        m.eval()
    return enc, lr, dec, ns


def state_dict_of(enc, lr, dec, ns):
    from collections import OrderedDict
    sd = OrderedDict()
    for prefix, m in (("encoders.", enc), ("length_regulator.", lr), ("decoders.", dec), ("noise_sampler.", ns)):
        for k, v in m.state_dict().items():
            sd[prefix + k] = v.detach().clone()
    return sd


def ref_predict_step(models, config, data, *, dwell_mean, dwell_std, noise_std, noise_sampling,
                     duration_sampling, min_noise, min_duration):
    """Transcription of model.py:195-240 around the reference modules."""
    enc, lr, dec, ns = models
    with torch.inference_mode():
        bs, seq_l = data.shape[:2]
        data = data.reshape(bs, seq_l, -1)
        enc_out, emb_out = enc(data)
        sigma = ns(emb_out)[:, :, None]
        lr_out, dur_f, _, sigma_ext, _ = lr(
            emb_out=emb_out, x=enc_out, target=None, noise_std_prediction=sigma,
            max_length=config["max_signal_len"], dwell_mean=dwell_mean, dwell_std=dwell_std,
            duration_sampling=duration_sampling, min_length=min_duration)
        p = dec(lr_out, None)
        pred = (p * config["scaling_max_value"]).squeeze(-1)
        if noise_std > 0:
            nz = pred != 0
            if noise_sampling:
                s = torch.clamp(sigma_ext, min=min_noise).squeeze(-1) * noise_std * config["scaling_max_value"]
                g = torch.normal(mean=0, std=s)
                pred[nz] += g[nz]
            else:
                g = torch.normal(mean=0, std=noise_std, size=pred.shape)
                pred[nz] += g[nz]
        pred = torch.clamp(pred, min=0)
    return dict(emb_out=emb_out, enc_out=enc_out, sigma=sigma.squeeze(-1), dur_f=dur_f,
                dur_i=torch.round(dur_f).int(), lr_out=lr_out, sigma_ext=sigma_ext.squeeze(-1),
                p=p.squeeze(-1), pA=pred)


def ref_export(read_ids, pred):
    """model.py:242-245, 262-286 (single flush, keep_last=False)."""
    from collections import defaultdict
    res = defaultdict(list)
    for rid, row in zip(read_ids, pred):
        res[rid].append(row)
    out = {}
    for k, v in res.items():
        cat = torch.cat(v)
        out[k] = cat[cat.nonzero()].squeeze()
    return out


def ref_digitise(signal, profile, rna):
    """signal_io.py:79-85, 134-141 verbatim arithmetic."""
    digitisation = float(profile["digitisation"])
    signal_range = float(profile["range"])
    offset = float(profile["offset_mean"])
    signal = signal.cpu().numpy().astype(np.float32)
    raw = np.round(signal * digitisation / signal_range - offset)
    raw = raw.astype(np.int16)
    if rna:
        raw = np.ascontiguousarray(raw[::-1])
    return raw


def codes_from_onehot(oh):
    """[.., k, 5] one-hot -> int8 letter codes, -1 for an all-zero row."""
    oh = np.asarray(oh, dtype=np.float32)
    code = oh.argmax(-1).astype(np.int8)
    code[oh.sum(-1) == 0] = -1
    return code


def read_fasta_plain(path):
    name, seq, out = None, [], []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if name is not None:
                    out.append((name, "".join(seq)))
                name, seq = line[1:].split()[0], []
            else:
                seq.append(line.strip())
    if name is not None:
        out.append((name, "".join(seq)))
    return out


def predict_fixture(ref_modules, ref_utils, models, config, reads, profile_name, tag, extra=None, **opts):
    profile = dict(ref_utils.get_profile(profile_name))
    ids, chunks = [], []
    for name, seq in reads:
        c = ref_utils.split_sequence(seq, config)
        if c.size > 0:
            ids += [name] * len(c)
            chunks.append(c)
    data_np = np.concatenate(chunks, 0)
    data = torch.from_numpy(data_np)
    torch.set_float32_matmul_precision("highest")
    st = ref_predict_step(models, config, data, **opts)
    torch.set_float32_matmul_precision("medium")          # what model.py:22 sets
    st_med = ref_predict_step(models, config, data, **opts)
    torch.set_float32_matmul_precision("highest")
    sig = ref_export(ids, st["pA"])
    rna = profile_name.startswith("rna")
    raws = [ref_digitise(s.reshape(-1), profile, rna) for s in sig.values()]
    pas = [s.reshape(-1).numpy() for s in sig.values()]
    offs = np.cumsum([0] + [len(r) for r in raws]).astype(np.int64)
    fx = dict(
        read_names=np.array([n for n, _ in reads]), read_seqs=np.array([s for _, s in reads]),
        chunk_read_names=np.array(ids), codes=codes_from_onehot(data_np),
        emb_out=st["emb_out"].numpy(), enc_out=st["enc_out"].numpy(), sigma=st["sigma"].numpy(),
        dur_f=st["dur_f"].numpy(), dur_i=st["dur_i"].numpy(), sigma_ext=st["sigma_ext"].numpy(),
        lr_out_first2=st["lr_out"][:2].numpy(), lr_out_sum=st["lr_out"].double().sum(-1).numpy(),
        p=st["p"].numpy(), pA=st["pA"].numpy(), p_medium=st_med["p"].numpy(),
        signal_names=np.array(list(sig.keys())), raw=np.concatenate(raws) if raws else np.zeros(0, np.int16),
        signal_pa=np.concatenate(pas) if pas else np.zeros(0, np.float32), raw_offsets=offs,
        profile=np.array(profile_name), opts=np.array(json.dumps(opts)),
    )
    if extra:
        fx.update(extra)
    np.savez_compressed(os.path.join(OUT, f"predict_{tag}.npz"), **fx)
    print(f"predict_{tag}: {len(reads)} reads, {data_np.shape[0]} chunks, {offs[-1]} samples")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_modules, ref_utils = import_reference()
    sys.path.insert(0, ROOT)
    from oracle import s2s_oracle as orc

    base_cfg = yaml.safe_load(open(os.path.join(REF_SRC, "seq2squiggle", "config.yaml")))
    rng = np.random.default_rng(20240917)

    test_reads = read_fasta_plain("/root/reference/example/test.fasta")
    synth = "".join(rng.choice(list("ACGT"), size=1003))
    reads_k9 = test_reads + [("synthetic-1003", synth)]

    # ---------------- k = 9 (dna-r10 / rna-004) ----------------
    cfg9 = dict(base_cfg)
    cfg9 = ref_utils.update_config("dna-r10-prom", cfg9)
    m9 = build_reference_model(ref_modules, cfg9, seed=1)
    sd9 = state_dict_of(*m9)
    torch.save(orc.lightning_checkpoint(sd9, cfg9), os.path.join(OUT, "ckpt_k9_seed1.ckpt"))

    ideal = dict(dwell_mean=5000 / 400, dwell_std=0.0, noise_std=0.0, noise_sampling=False,
                 duration_sampling=False, min_noise=0.0, min_duration=3)
    predict_fixture(ref_modules, ref_utils, m9, cfg9, reads_k9, "dna-r10-prom", "k9_ideal", **ideal)
    predict_fixture(ref_modules, ref_utils, m9, cfg9, reads_k9[:3], "rna-004-min", "k9_rna_ideal",
                    **dict(ideal, dwell_mean=4000 / 130))

    # sampler parameters (deterministic part of the duration / noise samplers) + one seeded draw
    with torch.inference_mode():
        data = torch.from_numpy(np.concatenate([ref_utils.split_sequence(s, cfg9) for _, s in reads_k9], 0))
        enc_out, emb_out = m9[0](data.reshape(data.shape[0], 16, -1))
        ds = m9[1].duration_sampler
        conc = torch.clamp(ds.conc_layer(emb_out), min=1e-8)
        rate = torch.clamp(ds.rate_layer(emb_out), min=1e-8)
        sigma = m9[3](emb_out)
        torch.manual_seed(123)
        dur_sample, _ = ds(emb_out)
    np.savez_compressed(os.path.join(OUT, "samplers_k9.npz"), codes=codes_from_onehot(data.numpy()),
                        conc=conc.squeeze(-1).numpy(), rate=rate.squeeze(-1).numpy(), sigma=sigma.numpy(),
                        dur_sample_seed123=dur_sample.numpy())

    # "biased" variant: out_linear.bias += 0.6 so almost every position is > 0 (trained-like range)
    with torch.no_grad():
        m9[2].out_linear.bias += 0.6
    predict_fixture(ref_modules, ref_utils, m9, cfg9, reads_k9, "dna-r10-prom", "k9_biased",
                    extra=dict(out_bias_delta=np.float32(0.6)), **ideal)
    with torch.no_grad():
        m9[2].out_linear.bias -= 0.6

    # ---------------- k = 6 (dna-r9) ----------------
    cfg6 = ref_utils.update_config("dna-r9-min", dict(base_cfg))
    m6 = build_reference_model(ref_modules, cfg6, seed=2)
    torch.save(orc.lightning_checkpoint(state_dict_of(*m6), cfg6), os.path.join(OUT, "ckpt_k6_seed2.ckpt"))
    odd = [("odd-letters", "ACGTNNacgtRYACGTACGGTTACAGGATTACCAGT_ACGTTGCAAGGTCCATG" * 3),
           ("short-5", "ACGTA"), ("exact-k", "ACGTAC"), ("len-21", "ACGTACGTTGCAACGTTAGCA"),
           ("synthetic-300", "".join(rng.choice(list("ACGT"), size=300)))]
    predict_fixture(ref_modules, ref_utils, m6, cfg6, odd, "dna-r9-min", "k6_ideal",
                    **dict(ideal, dwell_mean=4000 / 450))

    # ---------------- length-regulator KATs (modules.py:344-392) ----------------
    lr = m9[1]
    g = torch.Generator().manual_seed(7)
    x = torch.randn(6, 16, 64, generator=g)
    s = torch.rand(6, 16, 1, generator=g)
    dur = torch.tensor([
        [12] * 16,                                         # ideal R10: 192 < 250
        [31] * 16,                                         # RNA ideal: 496 -> cropped at 250
        [0, 3, 0, 0, 7, 1, 1, 0, 20, 0, 5, 5, 5, 0, 0, 2],  # zero durations skip the k-mer
        [250] + [0] * 15,                                  # one k-mer fills everything
        [1] * 16,                                          # 16 samples then zero fill
        [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 110],
    ], dtype=torch.int32)
    with torch.inference_mode():
        out, sext, _ = lr.LR(x, s, dur, max_length=250)
    np.savez_compressed(os.path.join(OUT, "lr_kat.npz"), x=x.numpy(), sigma=s.numpy(), dur=dur.numpy(),
                        out=out.numpy(), sigma_ext=sext.numpy())

    # ---------------- tokeniser KATs (utils.py:56-89, 334-356) ----------------
    tok = {}
    for k, cfg in ((9, cfg9), (6, cfg6)):
        for name, seq in [("plain40", "ACGTTGCAAGGTCCATGACGTTGCAAGGTCCATGACGTAC"), ("odd", odd[0][1][:60]),
                          ("short", "ACGT"), ("exact", "ACGTACGTA"[:k]), ("k_plus_15", synth[:k + 15]),
                          ("k_plus_16", synth[:k + 16]), ("underscore", "AC_GTACGTAC__ACGTTGCA")]:
            c = ref_utils.split_sequence(seq, cfg)
            tok[f"k{k}_{name}"] = dict(seq=seq, k=k, shape=list(c.shape),
                                       codes=codes_from_onehot(c).reshape(-1).tolist() if c.size else [])
    json.dump(tok, open(os.path.join(OUT, "tokeniser_kat.json"), "w"))

    # ---------------- digitise KATs (signal_io.py:134-141) ----------------
    pa = np.concatenate([rng.uniform(0, 200, 2000).astype(np.float32),
                         np.array([0.0, 1e-6, 165.0, 4500.0, 9000.0, 17.4999, 17.5, 18.5], np.float32)])
    dig = {}
    for pname in ("dna-r10-prom", "dna-r10-min", "dna-r9-min", "dna-r9-prom", "rna-004-min", "rna-004-prom"):
        prof = ref_utils.get_profile(pname)
        # also hit exact .5 boundaries for this profile: pA such that the pre-round value is h + 0.5
        half = (np.arange(-3, 60, dtype=np.float64) + 0.5 + prof["offset_mean"]) * prof["range"] / prof["digitisation"]
        arr = np.concatenate([pa, half[half > 0].astype(np.float32)])
        dig[pname + "/pa"] = arr
        dig[pname + "/raw"] = ref_digitise(torch.from_numpy(arr), prof, pname.startswith("rna"))
    np.savez_compressed(os.path.join(OUT, "digitise_kat.npz"), **dig)

    # ---------------- profiles + read sampler (utils.py:129-263, 311-331, 415-479) ----------------
    prof_all = {p: ref_utils.get_profile(p) for p in ("dna-r10-min", "dna-r10-prom", "dna-r9-min", "dna-r9-prom",
                                                       "rna-004-min", "rna-004-prom")}
    json.dump(prof_all, open(os.path.join(OUT, "profiles.json"), "w"), indent=1)

    genome = "".join(rng.choice(list("ACGTN"), size=30000, p=[0.245, 0.245, 0.245, 0.245, 0.02]))
    samp = dict(genome_seed_note="genome = rng(20240917) draw stored below", genome=genome, cases=[])
    for distr in ("expon", "gamma", "beta"):
        for profile in ("dna-r10-prom", "rna-004-min"):
            seed = 11
            random.seed(seed)
            reads = ref_utils.sampling(40, [genome], [len(genome)], 800, seed, len(genome), distr, profile, 30)
            samp["cases"].append(dict(distr=distr, profile=profile, seed=seed, n=40, r=800,
                                      lens=[len(r) for r in reads],
                                      md5=[hashlib.md5(r.encode()).hexdigest() for r in reads]))
    json.dump(samp, open(os.path.join(OUT, "read_sampling.json"), "w"))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
