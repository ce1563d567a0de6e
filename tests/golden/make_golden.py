#!/usr/bin/env python
"""Regenerates tests/golden/* from /root/reference (run in the authoring container).

Two kinds of vectors are written:

1. The reference's OWN golden data and known-answer vectors, repackaged as numpy
   arrays (values untouched):
     test/srfft_test.cc:11-271                    -> srfft128_in / srfft128_out
     test/data/en-us-hello.wav, en-us-cat.wav     -> hello_pcm / cat_pcm (int16 samples)
     test/data/fbankmat_en-us-hello.wav.txt       -> hello_fbank_kaldi [47][40]
     test/data/fbankcmvnmat_en-us-hello.wav.txt   -> hello_cmvn_kaldi  [47][40]
     test/data/cmvn_stats.bin                     -> cmvn_stats [41]
     test/nnet_test.cc:25-72                      -> nnet_kat_* (Linear / Softmax known answers)
2. Outputs of the UNMODIFIED reference compiled into oracle/_ref (oracle/build_ref.sh) on
   the repo wavs, on seeded synthetic PCM (pocketkaldi_b200.synth) and on the toy model
   written below with pocketkaldi_b200.formats: *_ref arrays, toy decode strings.

The GPU box has no /root/reference, so the -m gpu tests read these files.
"""

import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.oracle import Reference, build  # noqa: E402
from pocketkaldi_b200 import formats  # noqa: E402
from pocketkaldi_b200.synth import synth_pcm  # noqa: E402

REF = os.environ.get("PK_REFERENCE_DIR", "/root/reference")
TOY_WORDS = ["<eps>", "hello", "world", "cat", "milk"]


def parse_c_float_array(text, name):
    m = re.search(r"float\s+%s\[[^\]]*\]\s*=\s*\{(.*?)\};" % re.escape(name), text, re.S)
    vals = re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?(?=f?\s*[,}\n])", m.group(1) + "\n")
    return np.array([float(v) for v in vals], dtype=np.float32)


def toy_graph():
    """Word-loop grammar: 4 words x 3 emitting states with self-loops, so that the decoded
    word string depends on the acoustic scores (SURVEY.md Appendix B note)."""
    n_words, n_hmm = 4, 3
    arcs = []
    state_of = lambda w, s: 1 + w * n_hmm + s
    tid = 1
    tid2pdf = [0]
    enter_tid = {}
    for w in range(n_words):
        for s in range(n_hmm):
            pdf = w * n_hmm + s
            st = state_of(w, s)
            # self-loop
            arcs.append((st, st, tid, 0, 0.05))
            tid2pdf.append(pdf)
            tid += 1
            # forward tid that *enters* this state
            enter_tid[(w, s)] = tid
            tid2pdf.append(pdf)
            tid += 1
    for w in range(n_words):
        arcs.append((0, state_of(w, 0), enter_tid[(w, 0)], w + 1, 0.2))
        for s in range(1, n_hmm):
            arcs.append((state_of(w, s - 1), state_of(w, s), enter_tid[(w, s)], 0, 0.1))
        for w2 in range(n_words):
            arcs.append((state_of(w, n_hmm - 1), state_of(w2, 0), enter_tid[(w2, 0)], w2 + 1,
                         0.2 + 0.01 * w2))
    finals = {state_of(w, n_hmm - 1): 0.0 for w in range(n_words)}
    num_states = 1 + n_words * n_hmm
    return (num_states, 0, finals, arcs), tid2pdf, n_words * n_hmm


def write_toy_model(out_dir, cmvn_stats):
    rng = np.random.default_rng(7)
    fst, tid2pdf, num_pdfs = toy_graph()
    layers = formats.make_dnn(rng, 440, 64, 2, num_pdfs, normalize=True)
    # widen the output layer so that posteriors are peaky and frames disagree
    layers[-2] = ("linear", layers[-2][1] * 3.0, layers[-2][2])
    prior = rng.uniform(0.5, 1.5, num_pdfs).astype(np.float32)
    prior /= prior.sum()
    return formats.write_model_dir(out_dir, "toy", layers, prior, 5, 5, tid2pdf,
                                   cmvn_stats=cmvn_stats, fst=fst, words=TOY_WORDS)


def main():
    build()
    ref = Reference()
    out = {}

    # ---- 1. the reference's own vectors
    src = open(os.path.join(REF, "test/srfft_test.cc")).read()
    out["srfft128_out"] = parse_c_float_array(src, "fft_data")
    out["srfft128_in"] = parse_c_float_array(src, "data")
    assert out["srfft128_in"].size == 128 and out["srfft128_out"].size == 128
    for name in ("hello", "cat"):
        wav = os.path.join(REF, "test/data/en-us-%s.wav" % name)
        pcm = formats.read_wav16(wav)
        assert np.array_equal(ref.read_wav(wav), pcm.astype(np.float32))
        out[name + "_pcm"] = pcm
    out["hello_fbank_kaldi"] = np.loadtxt(
        os.path.join(REF, "test/data/fbankmat_en-us-hello.wav.txt"), dtype=np.float32).reshape(-1, 40)
    out["hello_cmvn_kaldi"] = np.loadtxt(
        os.path.join(REF, "test/data/fbankcmvnmat_en-us-hello.wav.txt"), dtype=np.float32).reshape(-1, 40)
    stats = formats.read_vector(os.path.join(REF, "test/data/cmvn_stats.bin"))
    assert stats.size == 41
    out["cmvn_stats"] = stats
    # test/nnet_test.cc:25-72 known answers
    out["nnet_kat_W"] = np.array([[0.1, 0.8, 0.9], [0.4, 0.2, 0.7], [0.2, 0.1, 0.1], [0.4, 0.3, 0.2]], np.float32)
    out["nnet_kat_b"] = np.array([0.1, -0.1, 0.2, -0.2], np.float32)
    out["nnet_kat_x"] = np.array([[0.3, -0.1, 0.9]], np.float32)
    out["nnet_kat_linear_y"] = np.array([[0.86, 0.63, 0.34, 0.07]], np.float32)
    out["nnet_kat_x4"] = np.array([[0.3, -0.1, 0.9, 0.2]], np.float32)
    out["nnet_kat_softmax_y"] = np.array([[0.2274135, 0.15243983, 0.41437442, 0.20577225]], np.float32)
    out["nnet_kat_relu_y"] = np.array([[0.3, 0.0, 0.9, 0.2]], np.float32)

    # ---- 2. compiled-reference outputs
    for name in ("hello", "cat"):
        fb = ref.fbank(out[name + "_pcm"].astype(np.float32))
        out[name + "_fbank_ref"] = fb
        out[name + "_cmvn_ref"] = ref.cmvn(fb, stats)
    # synthetic: utt 0 at 10 s (998 frames), utt 1 at 12.5 s (1248 frames > 2 CMVN windows),
    # and short edge cases
    for tag, utt, n in (("noise10", 0, 160000), ("noise12", 1, 200000), ("short400", 2, 400),
                        ("short559", 3, 559), ("short560", 4, 560)):
        pcm = synth_pcm(1234, [utt], n)[0]
        fb = ref.fbank(pcm.astype(np.float32))
        out[tag + "_fbank_ref"] = fb
        out[tag + "_cmvn_ref"] = ref.cmvn(fb, stats)
        out[tag + "_spec"] = np.array([1234, utt, n], np.int64)

    toy_dir = os.path.join(HERE, "toy")
    conf = write_toy_model(toy_dir, stats)
    am = ref.am_load(conf)
    for name in ("hello", "cat"):
        ll = ref.am_compute(am, out[name + "_cmvn_ref"])
        out[name + "_toy_loglik_ref"] = ll
    ll10 = ref.am_compute(am, out["noise10_cmvn_ref"])
    out["noise10_toy_loglik_ref"] = ll10
    tids = np.arange(1, 25, dtype=np.int32)
    dec, last = ref.decodable_fill(am, 0.1, out["hello_cmvn_ref"], tids)
    out["hello_toy_decodable_ref"] = dec
    out["hello_toy_islast_ref"] = last
    ref.am_free(am)

    hyps = {}
    for name in ("hello", "cat"):
        wav = os.path.join(REF, "test/data/en-us-%s.wav" % name)
        hyp, llpf = ref.decode_wav(conf, wav)
        hyps[name] = (hyp, llpf)
        out[name + "_toy_llpf_ref"] = np.float32(llpf)
    with open(os.path.join(HERE, "toy_decode_ref.txt"), "w") as fd:
        for name, (hyp, llpf) in hyps.items():
            fd.write("%s\t%s\t%.6f\n" % (name, hyp, llpf))

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    for k, v in sorted(out.items()):
        print("%-28s %s %s" % (k, v.dtype, v.shape))
    print(hyps)


if __name__ == "__main__":
    main()
