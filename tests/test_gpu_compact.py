"""Compact log-likelihood output (second half): fp16(t - off[frame]) + FP32 offset
per frame instead of the FP32 matrix, finished by the consumer. Checked against the FP32 output
of the same model, against the oracle, and through the host expansion helper of the C ABI."""

import numpy as np
import pytest

import pocketkaldi_b200 as pk
from pocketkaldi_b200 import formats
from pocketkaldi_b200.synth import synth_global_cmvn, synth_pcm

pytestmark = pytest.mark.gpu

LL_TOL = 2e-2  # north_star bar on the unscaled log-likelihood; the compact form must stay inside it


@pytest.fixture(scope="module")
def ctx():
    c = pk.Context(0)
    yield c
    c.close()


def both_forms(ctx, am, lens, g, scale):
    b = pk.Batch(ctx, lens, g, am, prob_scale=scale)
    b.synth_pcm(77, 0)
    b.run(pk.STAGE_ALL)
    full = b.get(pk.BUF_LOGLIK)
    b.set_compact(True)
    b.run(pk.STAGE_NNET)
    h = b.get(pk.BUF_LOGLIK16)
    off = b.get(pk.BUF_LOGLIK_OFF)
    with pytest.raises(pk.PkbError):
        b.get(pk.BUF_LOGLIK)
    exp = b.expand_compact(h, off, scale)
    b.set_compact(False)
    b.run(pk.STAGE_NNET)
    again = b.get(pk.BUF_LOGLIK)
    b.close()
    return full, h, off, exp, again


@pytest.mark.parametrize("pdfs,hidden,prior_kind", [(3000, 1024, "uniform"), (1000, 256, "skewed"), (1001, 128, "skewed"),
                                                    (200, 128, "uniform")])
def test_compact_matches_fp32_output(ctx, pdfs, hidden, prior_kind):
    rng = np.random.default_rng(pdfs)
    layers = formats.make_dnn(rng, 440, hidden, 2, pdfs)
    if prior_kind == "uniform":
        prior = np.full(pdfs, 1.0 / pdfs, np.float32)
    else:
        prior = rng.uniform(0.05, 3.0, pdfs).astype(np.float32)
        prior /= prior.sum()
    am = pk.AcousticModel(ctx, pk.PREC_BF16X3).from_layers(layers, prior, 5, 5)
    g = synth_global_cmvn()
    lens = [16000, 0, 48000, 399, 9000, 32000]   # ragged, with empty utterances
    scale = 0.1
    full, h, off, exp, again = both_forms(ctx, am, lens, g, scale)
    am.close()
    assert h.dtype == np.uint16 and h.shape == full.shape and off.shape == (full.shape[0],)
    assert np.array_equal(full, again)            # switching back restores the FP32 form bit for bit
    # numpy's own half decoding agrees with the library's host helper
    ref_exp = (h.view(np.float16).astype(np.float32) + off[:, None]) * np.float32(scale)
    assert np.array_equal(exp, ref_exp)
    err = np.abs(exp - full) / scale
    assert err.max() <= 1e-2 < LL_TOL             # <= 2^-7 for values within 32 of the frame's best
    # exact at the top of every frame, monotone below it: the argmax never moves
    top = full.argmax(1)
    assert np.array_equal(exp.argmax(1), top)
    assert np.max(err[np.arange(len(top)), top]) <= 1e-6
    # every stored value is <= 0 up to FP32 rounding of the two subtraction orders (the offset is the frame maximum)
    assert np.all(h.view(np.float16) <= 1e-4)


def test_compact_vs_float64_and_floor(ctx):
    # frames whose softmax underflows 1e-20 for most pdfs: the floor of src/am.cc:106-112 binds.
    # The reference's softmax has no max subtraction and overflows on such logits
    # (src/vector.cc:264-277), so the check is a float64 evaluation of the same formula.
    rng = np.random.default_rng(5)
    layers = formats.make_dnn(rng, 440, 64, 1, 256)
    layers[-2] = ("linear", (layers[-2][1] * 40.0).astype(np.float32), layers[-2][2])  # sharp logits
    prior = rng.uniform(0.5, 1.5, 256).astype(np.float32)
    prior /= prior.sum()
    am = pk.AcousticModel(ctx, pk.PREC_BF16X3).from_layers(layers, prior, 5, 5)
    g = synth_global_cmvn()
    pcm = synth_pcm(77, [0], 32000)[0]
    b = pk.Batch(ctx, [32000], g, am, prob_scale=1.0)
    b.set_pcm(pcm)
    b.run(pk.STAGE_ALL)
    feats = b.get(pk.BUF_FEATS).astype(np.float64)
    b.set_compact(True)
    b.run(pk.STAGE_NNET)
    exp = b.expand_compact(b.get(pk.BUF_LOGLIK16), b.get(pk.BUF_LOGLIK_OFF), 1.0)
    b.close()
    am.close()
    T = feats.shape[0]
    idx = np.clip(np.arange(T)[:, None] + np.arange(-5, 6)[None, :], 0, T - 1)
    x = feats[idx].reshape(T, 440)
    h = np.maximum(x @ layers[0][1].astype(np.float64).T + layers[0][2], 0.0)
    z = h @ layers[2][1].astype(np.float64).T + layers[2][2]
    lsm = z - (z.max(1, keepdims=True) + np.log(np.exp(z - z.max(1, keepdims=True)).sum(1, keepdims=True)))
    floor = np.log(np.float64(np.float32(1e-20)))
    ref = np.maximum(lsm, floor) - np.log(prior.astype(np.float64))[None, :]
    assert (lsm <= floor).mean() > 0.2                    # the floor really binds in this fixture
    d = np.abs(exp - ref)
    near = ref >= ref.max(1, keepdims=True) - 30.0
    assert d[near].max() <= LL_TOL
    assert d.max() <= 3.2e-2    # ~46 below the top: fp16 spacing 2^-5 -> at most 2^-6 + GEMM error
    assert np.mean(exp.argmax(1) == ref.argmax(1)) >= 0.999
