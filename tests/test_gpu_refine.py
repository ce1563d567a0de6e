"""PKB_PREC_FP16R: one FP16 MMA per product for every frame, then a second pass with the FP16C8
operands over the frames whose two best pdfs are nearly tied (include/pkb200.h).

The second pass gathers rows, so a refined frame must come out bit-identical to the same frame of
a pure FP16C8 run (GEMM rows are independent of each other); an unrefined frame carries the FP16
error, which has to stay inside the log-likelihood bar and cannot move its argmax.
"""

import numpy as np
import pytest

import pocketkaldi_b200 as pk
from pocketkaldi_b200 import formats
from pocketkaldi_b200.synth import synth_global_cmvn

pytestmark = pytest.mark.gpu

LL_TOL = 2e-2


@pytest.fixture(scope="module")
def ctx():
    c = pk.Context(0)
    yield c
    c.close()


def run(ctx, am, lens, g, compact, scale=1.0):
    b = pk.Batch(ctx, lens, g, am, prob_scale=scale)
    b.synth_pcm(31, 0)
    if compact:
        b.set_compact(True)
        b.run(pk.STAGE_ALL)
        out = b.expand_compact(b.get(pk.BUF_LOGLIK16), b.get(pk.BUF_LOGLIK_OFF), scale)
    else:
        b.run(pk.STAGE_ALL)
        out = b.get(pk.BUF_LOGLIK)
    stats = b.refine_stats()
    b.close()
    return out, stats


def top2_margin(ll):
    part = np.partition(ll, ll.shape[1] - 2, axis=1)
    return part[:, -1] - part[:, -2]


@pytest.mark.parametrize("compact", [False, True])
@pytest.mark.parametrize("pdfs,hidden,depth", [(3000, 512, 3), (1001, 128, 2)])
def test_refined_frames_equal_fp16c8_and_the_rest_stay_inside_the_bar(ctx, compact, pdfs, hidden, depth):
    rng = np.random.default_rng(pdfs + depth)
    layers = formats.make_dnn(rng, 440, hidden, depth, pdfs)
    prior = rng.uniform(0.2, 2.0, pdfs).astype(np.float32)
    prior /= prior.sum()
    g = synth_global_cmvn()
    lens = [48000, 0, 16000, 399, 160000, 9000]
    margin = 0.04
    am8 = pk.AcousticModel(ctx, pk.PREC_FP16C8).from_layers(layers, prior, 5, 5)
    ref, st8 = run(ctx, am8, lens, g, compact)
    am8.close()
    assert st8 == (0, 0)  # not a refining precision
    amr = pk.AcousticModel(ctx, pk.PREC_FP16R).from_layers(layers, prior, 5, 5)
    amr.set_refine_margin(margin)
    out, (rows, refined) = run(ctx, amr, lens, g, compact)
    frames = out.shape[0]
    assert out.shape == ref.shape and rows >= frames and 0 < refined < frames
    same = np.all(out == ref, axis=1)
    # every recomputed frame is bit-identical to the FP16C8 run (a few unrefined ones may be too)
    assert same.sum() >= refined
    # the FP16 pass's own error, seen where the 16-bit compact form adds none (within 1 of the
    # frame's best value); over the whole matrix the compact form may round the two runs to
    # different neighbours (2^-6 steps at distance 16..32 from the best value)
    near_top = ref >= ref.max(axis=1, keepdims=True) - 1.0
    err = np.max(np.where(near_top, np.abs(out - ref), 0.0), axis=1)
    assert err.max() <= LL_TOL / 2
    assert np.max(np.abs(out - ref)) <= (LL_TOL if compact else LL_TOL / 2)
    # a frame that was left alone has a clear winner, and the same one
    m = top2_margin(ref)
    assert np.all(m[~same] > margin - 2.5 * err.max())
    assert np.array_equal(out[~same].argmax(1), ref[~same].argmax(1))
    # frames with a near-tie in the FP16C8 output were all recomputed
    assert np.all(same[m < margin - 2.5 * err.max()])
    print("frames %d refined %d (%.1f %%), FP16 pass max |dLL| %.2e" % (frames, refined, 100.0 * refined / frames,
                                                                       err.max()))

    # margin 0: (almost) nothing is recomputed; a huge margin: everything is, bit for bit FP16C8
    amr.set_refine_margin(0.0)
    out0, (_, refined0) = run(ctx, amr, lens, g, compact)
    assert refined0 <= frames // 100
    assert np.max(np.abs(out0 - ref)) <= (LL_TOL if compact else LL_TOL / 2)
    amr.set_refine_margin(1000.0)
    out1, (_, refined1) = run(ctx, amr, lens, g, compact)
    amr.close()
    assert refined1 == frames
    assert np.array_equal(out1, ref)


def test_am_compute_and_other_outputs(ctx):
    # AcousticModel::Compute takes the refined path too; Nnet::Propagate (probabilities) runs FP16C8
    rng = np.random.default_rng(9)
    layers = formats.make_dnn(rng, 440, 256, 2, 700)
    prior = np.full(700, 1.0 / 700, np.float32)
    feats = [rng.standard_normal((t, 40)).astype(np.float32) for t in (57, 1, 300)]
    am8 = pk.AcousticModel(ctx, pk.PREC_FP16C8).from_layers(layers, prior, 5, 5)
    amr = pk.AcousticModel(ctx, pk.PREC_FP16R).from_layers(layers, prior, 5, 5)
    amr.set_refine_margin(1000.0)
    for a, b in zip(am8.compute_batch(feats), amr.compute_batch(feats)):
        assert np.array_equal(a, b)
    amr.set_refine_margin(0.04)
    for a, b in zip(am8.compute_batch(feats), amr.compute_batch(feats)):
        assert np.max(np.abs(a - b)) <= LL_TOL / 2
    am8.close()
    amr.close()
    x = rng.standard_normal((130, 440)).astype(np.float32)
    p8 = pk.Nnet(ctx, pk.PREC_FP16C8).from_layers(layers).Propagate(x)
    pr = pk.Nnet(ctx, pk.PREC_FP16R).from_layers(layers).Propagate(x)
    assert np.array_equal(p8, pr)


def test_margin_validation(ctx):
    layers = formats.make_dnn(np.random.default_rng(1), 440, 64, 1, 100)
    am = pk.AcousticModel(ctx, pk.PREC_FP16R).from_layers(layers, np.full(100, 0.01, np.float32), 5, 5)
    with pytest.raises(pk.PkbError):
        am.set_refine_margin(-1.0)
    am.close()


def test_streams_run_the_fp16c8_stages(ctx):
    # a stream's GEMMs have a few hundred rows: no selection pass, the FP16R model simply runs
    # its FP16C8 stages there, bit for bit what a FP16C8 model produces
    rng = np.random.default_rng(21)
    layers = formats.make_dnn(rng, 440, 256, 2, 600)
    prior = np.full(600, 1.0 / 600, np.float32)
    g = synth_global_cmvn()
    from pocketkaldi_b200.synth import synth_pcm
    S, chunk, n_chunks = 3, 2560, 6
    pcm = synth_pcm(99, np.arange(S), chunk * n_chunks)
    outs = {}
    for name, prec in (("c8", pk.PREC_FP16C8), ("r", pk.PREC_FP16R)):
        am = pk.AcousticModel(ctx, prec).from_layers(layers, prior, 5, 5)
        st = pk.Stream(ctx, am, S, chunk, g, 0.1)
        got = [st.push(pcm[:, k * chunk:(k + 1) * chunk].copy()) for k in range(n_chunks)]
        got.append(st.flush())
        outs[name] = [np.concatenate([np.asarray(o[s]).copy() for o in got]) for s in range(S)]
        st.close()
        am.close()
    for a, b in zip(outs["c8"], outs["r"]):
        assert a.shape == b.shape and a.shape[0] > 0
        assert np.array_equal(a, b)
