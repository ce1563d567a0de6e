"""Streaming (BASELINE config 5 shape): chunks pushed through pkb_stream_* with carried state
must reproduce the whole-utterance path; the reference has no streaming API, so the oracle is
the (reference-checked) batch path plus the compiled-reference golden log-likelihoods."""

import numpy as np
import pytest

import pocketkaldi_b200 as pk
from pocketkaldi_b200.synth import synth_pcm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = pk.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("chunk,n_chunks", [(2560, 7), (160, 40), (16000, 9)])
def test_stream_equals_whole_utterance(ctx, golden, toy_conf, chunk, n_chunks):
    am = pk.AcousticModel(ctx, pk.PREC_BF16X3).Read(toy_conf)
    S = 5
    total = chunk * n_chunks
    pcm = synth_pcm(1234, np.arange(S), total)
    whole = am.pcm_to_loglik([pcm[s] for s in range(S)], golden["cmvn_stats"], 0.1)
    st = pk.Stream(ctx, am, S, chunk, golden["cmvn_stats"], 0.1)
    got = [[] for _ in range(S)]
    for k in range(n_chunks):
        o = st.push(pcm[:, k * chunk:(k + 1) * chunk])
        for s in range(S):
            got[s].append(o[s].copy())
    o = st.flush()
    for s in range(S):
        got[s].append(o[s].copy())
        cat = np.concatenate(got[s])
        assert cat.shape == whole[s].shape
        assert np.max(np.abs(cat - whole[s])) < 1e-5 if cat.size else True
    # a second utterance through the same stream object (state was reset by flush)
    pcm2 = synth_pcm(99, np.arange(S), chunk * 3)
    outs = [st.push(pcm2[:, k * chunk:(k + 1) * chunk]).copy() for k in range(3)] + [st.flush().copy()]
    whole2 = am.pcm_to_loglik([pcm2[s] for s in range(S)], golden["cmvn_stats"], 0.1)
    for s in range(S):
        cat = np.concatenate([o[s] for o in outs])
        assert cat.shape == whole2[s].shape
        if cat.size:
            assert np.max(np.abs(cat - whole2[s])) < 1e-5
    st.close()


def test_stream_long_utterance_crosses_cmvn_window(ctx, golden, toy_conf):
    # 12.5 s > 600 frames: the ring-buffer subtraction must match the batch recurrence, which
    # is bit-exact against the reference (noise12 golden)
    am = pk.AcousticModel(ctx, pk.PREC_BF16X3).Read(toy_conf)
    seed, utt, n = [int(v) for v in golden["noise12_spec"]]
    pcm = synth_pcm(seed, [utt], n)
    chunk = 4000
    st = pk.Stream(ctx, am, 1, chunk, golden["cmvn_stats"], 1.0)
    outs = [st.push(pcm[:, k * chunk:(k + 1) * chunk]).copy() for k in range(n // chunk)]
    outs.append(st.flush().copy())
    cat = np.concatenate([o[0] for o in outs])
    whole = am.pcm_to_loglik([pcm[0]], golden["cmvn_stats"], 1.0)[0]
    assert cat.shape == whole.shape == (1248, 12)
    assert np.max(np.abs(cat - whole)) < 1e-5
    st.close()


def test_stream_graph_replay_with_stable_buffers(ctx, golden, toy_conf):
    # the same host buffers on every push: from the third chunk on the captured CUDA graph of the
    # steady-state launch sequence is replayed; results must still equal the whole-utterance path,
    # also across a flush and a second utterance, and for the compact output form
    am = pk.AcousticModel(ctx, pk.PREC_FP16C8).Read(toy_conf)
    S, chunk, n_chunks = 7, 2560, 25
    pcm = synth_pcm(4321, np.arange(S), chunk * n_chunks)
    whole = am.pcm_to_loglik([pcm[s] for s in range(S)], golden["cmvn_stats"], 0.1)
    st = pk.Stream(ctx, am, S, chunk, golden["cmvn_stats"], 0.1)
    buf = np.empty((S, chunk), np.int16)
    for rep in range(2):
        got = [[] for _ in range(S)]
        for k in range(n_chunks):
            buf[:] = pcm[:, k * chunk:(k + 1) * chunk]
            o = st.push(buf)
            for s in range(S):
                got[s].append(o[s].copy())
        o = st.flush()
        for s in range(S):
            cat = np.concatenate(got[s] + [o[s].copy()])
            assert cat.shape == whole[s].shape
            assert np.max(np.abs(cat - whole[s])) < 1e-5
    launches = ctx.profile_get()
    assert launches["gemm"][0] > 0 and launches["gemm_final"][0] >= 2 * n_chunks  # replays are counted
    # compact form through the same stream object
    st.set_compact(True)
    got = [[] for _ in range(S)]
    for k in range(n_chunks):
        buf[:] = pcm[:, k * chunk:(k + 1) * chunk]
        h, off = st.push_compact(buf)
        for s in range(S):
            got[s].append((h[s].view(np.float16).astype(np.float32) + off[s][:, None]) * np.float32(0.1))
    h, off = st.flush_compact()
    for s in range(S):
        cat = np.concatenate(got[s] + [(h[s].view(np.float16).astype(np.float32) + off[s][:, None]) * np.float32(0.1)])
        assert cat.shape == whole[s].shape
        assert np.max(np.abs(cat - whole[s])) < 1e-3          # 0.1 * 2^-7 for values within 32 of the best
        assert np.array_equal(cat.argmax(1), whole[s].argmax(1))
    with pytest.raises(pk.PkbError):
        st.push(buf)
    st.close()
    am.close()
