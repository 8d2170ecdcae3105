"""The CPU oracle against the golden vectors produced by the reference's own modules
(oracle/make_golden.py).  Integer results are bit-exact; fp32 tensors are compared with a
1e-5 absolute bound (identical ATen kernels give 0.0 on the generating host; the bound only
allows for a different CPU's GEMM summation order)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import s2s_oracle as orc

FP32_ATOL = 1e-5


def load_ckpt(golden_dir, name):
    ck = torch.load(os.path.join(golden_dir, name), map_location="cpu", weights_only=False)
    return ck["state_dict"], ck["hyper_parameters"]["config"]


def onehot_from_codes(codes):
    codes = np.asarray(codes)
    oh = np.zeros(codes.shape + (5,), dtype=np.float16)
    idx = np.nonzero(codes >= 0)
    oh[idx + (codes[idx],)] = 1
    return oh


@pytest.mark.parametrize("tag,ckpt", [("k9_ideal", "ckpt_k9_seed1.ckpt"), ("k9_rna_ideal", "ckpt_k9_seed1.ckpt"),
                                      ("k9_biased", "ckpt_k9_seed1.ckpt"), ("k6_ideal", "ckpt_k6_seed2.ckpt")])
def test_predict_stages_match_reference(golden_dir, tag, ckpt):
    sd, cfg = load_ckpt(golden_dir, ckpt)
    fx = np.load(os.path.join(golden_dir, f"predict_{tag}.npz"), allow_pickle=False)
    if "out_bias_delta" in fx.files:
        sd = dict(sd)
        sd["decoders.out_linear.bias"] = sd["decoders.out_linear.bias"] + float(fx["out_bias_delta"])
    opts = json.loads(str(fx["opts"]))
    # tokeniser: reads -> chunks must reproduce the reference's one-hot exactly
    chunks, ids = [], []
    for name, seq in zip(fx["read_names"], fx["read_seqs"]):
        c = orc.split_sequence(str(seq), cfg)
        assert np.array_equal(c, orc.split_sequence_fast(str(seq), cfg))
        if c.size:
            chunks.append(c)
            ids += [str(name)] * len(c)
    data = np.concatenate(chunks, 0)
    assert np.array_equal(data, onehot_from_codes(fx["codes"]))
    assert ids == [str(x) for x in fx["chunk_read_names"]]

    pred, st = orc.predict_step(sd, cfg, torch.from_numpy(data), return_stages=True, **opts)
    assert np.array_equal(st["dur_i"].numpy(), fx["dur_i"])
    for key in ("emb_out", "enc_out", "sigma", "sigma_ext", "p"):
        np.testing.assert_allclose(st[key].numpy(), fx[key], rtol=0, atol=FP32_ATOL, err_msg=key)
    np.testing.assert_allclose(st["lr_out"][:2].numpy(), fx["lr_out_first2"], rtol=0, atol=FP32_ATOL)
    np.testing.assert_allclose(pred.numpy(), fx["pA"], rtol=0, atol=FP32_ATOL * 165)

    # integer stages fed with the reference's own pA: assembly (zero strip) + digitisation are bit-exact
    sig = orc.assemble_reads(ids, torch.from_numpy(fx["pA"]))
    assert list(sig.keys()) == [str(x) for x in fx["signal_names"]]
    from oracle.profiles_kat import PROFILES
    prof = PROFILES[str(fx["profile"])]
    raws = [orc.digitise(s.reshape(-1).numpy(), prof["digitisation"], prof["range"], prof["offset_mean"],
                         rna=str(fx["profile"]).startswith("rna")) for s in sig.values()]
    assert np.array_equal(np.cumsum([0] + [len(r) for r in raws]), fx["raw_offsets"])
    assert np.array_equal(np.concatenate(raws), fx["raw"])


def test_known_answer_integers(golden_dir):
    """SURVEY §8c integer facts: chunk counts, ideal dwell, half-even rounding."""
    assert orc.n_chunks_of_read(1000, 9) == 62
    assert orc.n_chunks_of_read(96, 9) == 6 and orc.n_chunks_of_read(104, 9) == 6
    assert orc.n_chunks_of_read(8, 9) == 0 and orc.n_chunks_of_read(9, 9) == 1
    assert int(torch.round(torch.tensor(12.5))) == 12 and int(torch.round(torch.tensor(13.5))) == 14
    fx = np.load(os.path.join(golden_dir, "predict_k9_ideal.npz"))
    assert (fx["dur_i"] == 12).all()
    fx6 = np.load(os.path.join(golden_dir, "predict_k6_ideal.npz"))
    assert (fx6["dur_i"] == 9).all()


def test_length_regulator_kat(golden_dir):
    fx = np.load(os.path.join(golden_dir, "lr_kat.npz"))
    out, sext = orc.lr_expand(torch.from_numpy(fx["x"]), torch.from_numpy(fx["sigma"]),
                              torch.from_numpy(fx["dur"]), 250)
    assert np.array_equal(out.numpy(), fx["out"]) and np.array_equal(sext.numpy(), fx["sigma_ext"])
    # integer-index form == the reference's alignment-matrix bmm (exact copy of rows / zero fill)
    j = orc.lr_expand_indices(fx["dur"], 250)
    gathered = np.where(j[..., None] >= 0, np.take_along_axis(fx["x"], np.maximum(j, 0)[..., None], axis=1), 0.0)
    assert np.array_equal(gathered.astype(np.float32), fx["out"])
    for b in range(fx["dur"].shape[0]):
        ri = np.repeat(np.arange(16), fx["dur"][b])[:250]
        assert np.array_equal(j[b, :len(ri)], ri) and (j[b, len(ri):] == -1).all()


def test_tokeniser_kat(golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "tokeniser_kat.json")))
    for name, case in kat.items():
        cfg = dict(orc.DEFAULT_CONFIG, seq_kmer=case["k"])
        for fn in (orc.split_sequence, orc.split_sequence_fast):
            c = fn(case["seq"], cfg)
            if not case["codes"]:           # reference returns an empty (0,) array for reads shorter than k
                assert c.size == 0, name
                continue
            assert list(c.shape) == case["shape"], name
            if c.size:
                code = c.astype(np.float32).argmax(-1).astype(np.int8)
                code[c.sum(-1) == 0] = -1
                assert code.reshape(-1).tolist() == case["codes"], name


def test_digitise_kat(golden_dir):
    fx = np.load(os.path.join(golden_dir, "digitise_kat.npz"))
    from oracle.profiles_kat import PROFILES
    for pname, prof in PROFILES.items():
        raw = orc.digitise(fx[pname + "/pa"], prof["digitisation"], prof["range"], prof["offset_mean"],
                           rna=pname.startswith("rna"))
        assert np.array_equal(raw, fx[pname + "/raw"]), pname


def test_sampler_parameters_and_seeded_draw(golden_dir):
    sd, cfg = load_ckpt(golden_dir, "ckpt_k9_seed1.ckpt")
    fx = np.load(os.path.join(golden_dir, "samplers_k9.npz"))
    data = torch.from_numpy(onehot_from_codes(fx["codes"]))
    _, emb = orc.encoder_forward(sd, cfg, data.reshape(data.shape[0], 16, -1))
    conc, rate = orc.duration_params(sd, emb)
    np.testing.assert_allclose(conc.squeeze(-1).numpy(), fx["conc"], rtol=0, atol=FP32_ATOL)
    np.testing.assert_allclose(rate.squeeze(-1).numpy(), fx["rate"], rtol=0, atol=FP32_ATOL)
    np.testing.assert_allclose(orc.noise_sampler_forward(sd, emb).numpy(), fx["sigma"], rtol=0, atol=FP32_ATOL)
    torch.manual_seed(123)
    d = orc.duration_sampler_forward(sd, emb)
    np.testing.assert_allclose(d.numpy(), fx["dur_sample_seed123"], rtol=1e-4, atol=1e-4)


def test_random_init_matches_reference_checkpoint(golden_dir):
    """Our reference-free checkpoint generator reproduces the reference model's init stream."""
    for name, seed in (("ckpt_k9_seed1.ckpt", 1), ("ckpt_k6_seed2.ckpt", 2)):
        sd, cfg = load_ckpt(golden_dir, name)
        mine = orc.random_init_state_dict(cfg, seed)
        assert list(mine.keys()) == list(sd.keys())
        assert len(sd) == 84
        for k in sd:
            assert torch.equal(mine[k], sd[k]), k
