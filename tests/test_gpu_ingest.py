"""SURVEY 8(f)-2 end to end on the GPU: .scp list -> validated headers -> int16 straight into
pinned staging -> device batch -> log-likelihoods, against the compiled-reference golden vectors
of the same wavs (which the reference reads one at a time through float)."""

import numpy as np
import pytest

import pocketkaldi_b200 as pk
from pocketkaldi_b200 import formats
from pocketkaldi_b200.binding import PinnedArray

pytestmark = pytest.mark.gpu


def test_scp_to_loglik_matches_reference_golden(tmp_path, golden, toy_conf):
    names = ["hello", "cat", "hello"]
    paths = []
    for i, n in enumerate(names):
        p = str(tmp_path / ("%d_%s.wav" % (i, n)))
        formats.write_wav16(p, golden[n + "_pcm"])
        paths.append(p)
    scp = str(tmp_path / "list.scp")
    open(scp, "w").write("\n".join(paths) + "\n")

    ctx = pk.Context(0)
    am = pk.AcousticModel(ctx, pk.PREC_BF16X3).Read(toy_conf)
    wl = pk.WavList(scp)
    batch = pk.Batch(ctx, wl.num_samples, golden["cmvn_stats"], am=am, prob_scale=1.0)
    staging = PinnedArray((batch.total_samples,), np.int16)
    wl.read_i16(out=staging.array, n_threads=3)
    batch.set_pcm(staging.array)
    batch.run()
    ll = batch.get(pk.BUF_LOGLIK)
    feats = batch.get(pk.BUF_FEATS)
    off = np.concatenate([[0], np.cumsum(batch.num_frames)])
    for i, n in enumerate(names):
        ref = golden[n + "_toy_loglik_ref"]
        got = ll[off[i]:off[i + 1]]
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) < 2e-2
        assert np.mean(got.argmax(1) == ref.argmax(1)) >= 0.999
        cref = golden[n + "_cmvn_ref"]
        assert np.max(np.abs(feats[off[i]:off[i + 1]] - cref) / np.maximum(1.0, np.abs(cref))) < 1e-4
    # utterances 0 and 2 are the same file: identical bits
    assert np.array_equal(ll[off[0]:off[1]], ll[off[2]:off[3]])
    staging.free()
    batch.close()
    wl.close()
    ctx.close()
