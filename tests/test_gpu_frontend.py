"""GPU parity of the front end (fbank + CMVN) through the C ABI, against the golden vectors
of the compiled reference, the restatement oracle, and size-independent properties."""

import numpy as np
import pytest

import pocketkaldi_b200 as pk
from pocketkaldi_b200.synth import synth_pcm

pytestmark = pytest.mark.gpu

FBANK_RTOL = 1e-4   # north_star: fbank within 1e-4 relative
CMVN_TOL = 1e-4     # |d| <= 1e-4 * max(1, |ref|)  (outputs cross zero)


@pytest.fixture(scope="module")
def ctx():
    c = pk.Context(0)
    yield c
    c.close()


def pcm_of(golden, name):
    if name + "_pcm" in golden:
        return golden[name + "_pcm"]
    seed, utt, n = golden[name + "_spec"]
    return synth_pcm(int(seed), [int(utt)], int(n))[0]


def fbank_err(got, ref):
    return float(np.max(np.abs(got - ref) / np.abs(ref))) if ref.size else 0.0


def cmvn_err(got, ref):
    return float(np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref)))) if ref.size else 0.0


@pytest.mark.parametrize("name", ["hello", "cat", "noise10", "noise12", "short400", "short559", "short560"])
def test_fbank_vs_reference_golden(ctx, golden, name):
    pcm = pcm_of(golden, name)
    ref = golden[name + "_fbank_ref"]
    got_f = pk.Fbank(ctx).Compute(pcm.astype(np.float32))   # reference call shape: float wave
    got_i = pk.Fbank(ctx).Compute(pcm)                      # int16 ingestion
    assert got_f.shape == ref.shape and got_i.shape == ref.shape
    assert fbank_err(got_f, ref) < FBANK_RTOL
    assert np.array_equal(got_f, got_i)


def test_fbank_vs_kaldi_golden_text(ctx, golden):
    # test/fbank_test.cc:15-56 against test/data/fbankmat_en-us-hello.wav.txt
    got = pk.Fbank(ctx).Compute(golden["hello_pcm"].astype(np.float32))
    assert np.max(np.abs(got - golden["hello_fbank_kaldi"])) < 5e-5


def test_fbank_vs_oracle_random(ctx, oracle):
    pcm = synth_pcm(77, [3], 32000)[0]
    ref = oracle.fbank(pcm.astype(np.float32))
    got = pk.Fbank(ctx).Compute(pcm)
    assert fbank_err(got, ref) < FBANK_RTOL
    # a speech-like signal with a large DC offset and strong spectral tilt
    t = np.arange(24000, dtype=np.float64)
    x = 8000 + 6000 * np.sin(2 * np.pi * 180 * t / 16000) + 30 * np.sin(2 * np.pi * 5200 * t / 16000)
    x = np.round(x).astype(np.float32)
    assert fbank_err(pk.Fbank(ctx).Compute(x), oracle.fbank(x)) < FBANK_RTOL


def test_fbank_batch_ragged_and_empty(ctx, golden, oracle):
    lens = [160000, 399, 0, 400, 8359, 1000, 47999]
    pcms = [synth_pcm(5, [i], n)[0] for i, n in enumerate(lens)]
    outs = ctx.fbank_batch(pcms)
    assert [o.shape[0] for o in outs] == [oracle.num_frames(n) for n in lens]
    for p, o in zip(pcms, outs):
        assert fbank_err(o, oracle.fbank(p.astype(np.float32))) < FBANK_RTOL
    assert ctx.fbank_batch([]) == []
    # batch result == one-by-one result, bit for bit (no cross-utterance coupling)
    for p, o in zip(pcms, outs):
        assert np.array_equal(pk.Fbank(ctx).Compute(p), o)


@pytest.mark.parametrize("name", ["hello", "cat", "noise10", "noise12", "short400", "short560"])
def test_cmvn_bit_exact_on_reference_raw(ctx, golden, name):
    # same raw input as the reference -> identical output, incl. > 600 frames (noise12)
    raw = golden[name + "_fbank_ref"]
    cm = pk.CMVN(ctx, golden["cmvn_stats"], raw)
    got = np.stack([cm.GetFrame(t) for t in range(raw.shape[0])]) if raw.shape[0] else raw
    assert np.array_equal(got, golden[name + "_cmvn_ref"])


def test_cmvn_vs_kaldi_golden_text(ctx, golden):
    # test/cmvn_test.cc:33-82 against test/data/fbankcmvnmat_en-us-hello.wav.txt
    raw = pk.Fbank(ctx).Compute(golden["hello_pcm"].astype(np.float32))
    got = ctx.cmvn_batch([raw], golden["cmvn_stats"])[0]
    assert np.max(np.abs(got - golden["hello_cmvn_kaldi"])) < 5e-5


def test_frontend_end_to_end_tolerance(ctx, golden):
    for name in ["hello", "cat", "noise10", "noise12"]:
        raw = pk.Fbank(ctx).Compute(pcm_of(golden, name))
        got = ctx.cmvn_batch([raw], golden["cmvn_stats"])[0]
        assert cmvn_err(got, golden[name + "_cmvn_ref"]) < CMVN_TOL


def test_cmvn_batch_ragged(ctx, golden, oracle):
    rng = np.random.default_rng(0)
    raws = [(rng.standard_normal((n, 40)) * 3 + 17).astype(np.float32) for n in (1, 0, 599, 600, 601, 1500)]
    outs = ctx.cmvn_batch(raws, golden["cmvn_stats"])
    for r, o in zip(raws, outs):
        assert np.array_equal(o, oracle.cmvn(r, golden["cmvn_stats"]))


def test_batch_pipeline_frontend_full_size_properties(ctx, golden):
    # config 2 shape: 360 x 10 s. Device-side synthetic PCM equals the numpy generator,
    # utterances are independent (permutation property), shifted copies agree frame-wise.
    n_utts, n = 360, 160000
    b = pk.Batch(ctx, [n] * n_utts, golden["cmvn_stats"])
    b.synth_pcm(1234, 0)
    b.run(pk.STAGE_FBANK | pk.STAGE_CMVN)
    pcm = b.get(pk.BUF_PCM).reshape(n_utts, n)
    assert np.array_equal(pcm[[0, 7, 359]], synth_pcm(1234, [0, 7, 359], n))
    raw = b.get(pk.BUF_RAW).reshape(n_utts, 998, 40)
    feats = b.get(pk.BUF_FEATS).reshape(n_utts, 998, 40)
    assert fbank_err(raw[0], golden["noise10_fbank_ref"]) < FBANK_RTOL
    assert cmvn_err(feats[0], golden["noise10_cmvn_ref"]) < CMVN_TOL
    assert np.isfinite(feats).all()
    # the same utterance ids at other batch positions give identical bits
    b2 = pk.Batch(ctx, [n] * 4, golden["cmvn_stats"])
    b2.synth_pcm(1234, 100)
    b2.run(pk.STAGE_FBANK | pk.STAGE_CMVN)
    assert np.array_equal(b2.get(pk.BUF_FEATS).reshape(4, 998, 40), feats[100:104])
    chk = b.checksum(pk.BUF_FEATS)
    assert abs(chk - float(feats.astype(np.float64).sum())) < 1e-3 * max(1.0, abs(chk))
    b.close()
    b2.close()
