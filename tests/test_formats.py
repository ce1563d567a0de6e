"""Host logic: model-file writers/readers round-trip and are accepted by the compiled reference."""

import numpy as np
import pytest

from pocketkaldi_b200 import formats
from pocketkaldi_b200.synth import synth_pcm


def test_nnet_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    layers = formats.make_dnn(rng, 33, 17, 2, 9, normalize=True)
    p = str(tmp_path / "a.nnet")
    formats.write_nnet(p, layers)
    back = formats.read_nnet(p)
    assert [l[0] for l in back] == [l[0] for l in layers]
    for a, b in zip(layers, back):
        if a[0] == "linear":
            assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_vector_roundtrip_and_corruption(tmp_path):
    p = str(tmp_path / "v.bin")
    formats.write_vector(p, np.arange(5, dtype=np.float32))
    assert np.array_equal(formats.read_vector(p), np.arange(5, dtype=np.float32))
    raw = bytearray(open(p, "rb").read())
    raw[4] ^= 0x7
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        formats.read_vector(p)


def test_wav_roundtrip_through_reference_reader(tmp_path, reference):
    pcm = synth_pcm(5, [1], 1234)[0]
    p = str(tmp_path / "x.wav")
    formats.write_wav16(p, pcm)
    assert np.array_equal(formats.read_wav16(p), pcm)
    if reference is not None:
        assert np.array_equal(reference.read_wav(p), pcm.astype(np.float32))


def test_synth_pcm_is_counter_based():
    a = synth_pcm(1234, [0, 1, 2, 3], 1000)
    b = synth_pcm(1234, [2], 1000)
    c = synth_pcm(1234, [2], 500)
    assert np.array_equal(a[2], b[0]) and np.array_equal(b[0][:500], c[0])
    assert a.dtype == np.int16 and abs(float(a.std()) - 3000.0) < 200.0
    assert not np.array_equal(synth_pcm(1235, [2], 1000), b)


def test_model_dir_loads_in_reference(tmp_path, reference, golden):
    if reference is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(1)
    layers = formats.make_dnn(rng, 440, 32, 1, 6)
    prior = np.full(6, 1 / 6, np.float32)
    conf = formats.write_model_dir(str(tmp_path), "m", layers, prior, 5, 5, [0, 0, 1, 2, 3, 4, 5],
                                   cmvn_stats=golden["cmvn_stats"])
    am = reference.am_load(conf)
    assert reference.am_num_pdfs(am) == 6 and reference.am_tid2pdf(am, 3) == 2
    reference.am_free(am)


def test_mul_layer_roundtrip_fold_and_reference_rejection(tmp_path, reference, oracle, golden):
    # the MUL layer tool/convert_am.py:86-110 writes for a FixedScaleComponent
    rng = np.random.default_rng(3)
    base = formats.make_dnn(rng, 440, 16, 1, 5)
    v0 = rng.uniform(0.5, 2.0, 440).astype(np.float32)
    v1 = rng.uniform(-2.0, 2.0, 16).astype(np.float32)
    layers = [("mul", v0), base[0], ("mul", v1)] + base[1:]
    p = str(tmp_path / "m.nnet")
    formats.write_nnet(p, layers)
    back = formats.read_nnet(p)
    assert [l[0] for l in back] == [l[0] for l in layers]
    assert np.array_equal(back[0][1], v0) and np.array_equal(back[2][1], v1)
    folded = formats.fold_mul_layers(layers)
    assert [l[0] for l in folded] == [l[0] for l in base]
    # folding is the same function: oracle(folded)(x) == oracle(base)(x * v0) with rows scaled by v1
    x = rng.normal(size=(7, 440)).astype(np.float32)
    W, b = base[0][1], base[0][2]
    want = (x * v0) @ W.T.astype(np.float64) + b
    got = x @ folded[0][1].T.astype(np.float64) + folded[0][2]
    assert np.allclose(got, want * v1, rtol=1e-5, atol=1e-5)
    if reference is not None:
        prior = np.full(5, 0.2, np.float32)
        conf = formats.write_model_dir(str(tmp_path), "mm", layers, prior, 5, 5, [0, 1, 2, 3, 4])
        with pytest.raises(Exception):
            reference.am_load(conf)       # src/nnet.cc:122-126 "unexpected layer type"
        conf2 = formats.write_model_dir(str(tmp_path), "mf", folded, prior, 5, 5, [0, 1, 2, 3, 4])
        reference.am_load(conf2)
