"""End-to-end decode demonstration (not a graded benchmark): the reference CLI against the batch CLI
of this repo on the same model directory and the same wave files.

  python tools/decode_demo.py --utts 16 [--ref] [--precision bf16x3]

Writes a config-3-sized model (splice +-5 -> 6x1024 ReLU -> 3000 pdfs, random init) with a word-loop
graph, N synthetic 10 s utterances and a .scp list into a scratch directory, then times
  oracle/_ref/pocketkaldi_ref       (unmodified reference: one file at a time, CPU nnet)      [--ref]
  oracle/_ref/pocketkaldi_b200_cli  (reference main/decoder, acoustic half on the GPU per utterance)
  oracle/_ref/pocketkaldi_b200_batch (list ingestion + one GPU batch + decoder on host threads)
  the same with --gpu-decode 1      (the Viterbi search on the GPU as well, pkb_batch_decode)
and reports wall-clock seconds and how many hypotheses agree with the first CLI that ran.
"""

import argparse
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pocketkaldi_b200 import formats  # noqa: E402
from pocketkaldi_b200.synth import synth_global_cmvn, synth_pcm  # noqa: E402


def word_loop_graph(n_words, n_hmm, num_pdfs):
    """n_words words of n_hmm emitting states with self-loops, any word after any word; the
    (word, state) pairs are spread over the pdf range."""
    arcs, tid2pdf, enter = [], [0], {}
    state_of = lambda w, s: 1 + w * n_hmm + s
    stride = max(1, num_pdfs // (n_words * n_hmm))
    tid = 1
    for w in range(n_words):
        for s in range(n_hmm):
            pdf = (w * n_hmm + s) * stride
            arcs.append((state_of(w, s), state_of(w, s), tid, 0, 0.05))
            tid2pdf.append(pdf)
            tid += 1
            enter[(w, s)] = tid
            tid2pdf.append(pdf)
            tid += 1
    for w in range(n_words):
        arcs.append((0, state_of(w, 0), enter[(w, 0)], w + 1, 0.2))
        for s in range(1, n_hmm):
            arcs.append((state_of(w, s - 1), state_of(w, s), enter[(w, s)], 0, 0.1))
        for w2 in range(n_words):
            arcs.append((state_of(w, n_hmm - 1), state_of(w2, 0), enter[(w2, 0)], w2 + 1, 0.2 + 0.001 * w2))
    finals = {state_of(w, n_hmm - 1): 0.0 for w in range(n_words)}
    return (1 + n_words * n_hmm, 0, finals, arcs), tid2pdf


def run(cli, conf, scp, env, extra=()):
    e = dict(os.environ)
    e.update(env)
    t0 = time.perf_counter()
    r = subprocess.run([cli, conf, scp] + list(extra), stdout=subprocess.PIPE,
                       stderr=None if os.environ.get("PKB_CLI_TIMING") else subprocess.DEVNULL, env=e, check=True)
    out = r.stdout.decode()
    dt = time.perf_counter() - t0
    hyps = [line.split("\t")[1].strip() for line in out.strip().splitlines()]
    return dt, hyps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=16)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--ref", action="store_true", help="also time the unmodified reference CLI")
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()

    rng = np.random.default_rng(0)
    layers = formats.make_dnn(rng, 440, 1024, 6, 3000)
    prior = np.full(3000, 1.0 / 3000, np.float32)
    fst, tid2pdf = word_loop_graph(40, 3, 3000)
    words = ["<eps>"] + ["w%02d" % i for i in range(40)]
    with tempfile.TemporaryDirectory() as d:
        conf = formats.write_model_dir(d, "demo", layers, prior, 5, 5, tid2pdf,
                                       cmvn_stats=synth_global_cmvn(), fst=fst, words=words)
        n = int(args.seconds * 16000)
        pcm = synth_pcm(1234, list(range(args.utts)), n)
        paths = []
        for u in range(args.utts):
            p = os.path.join(d, "utt%04d.wav" % u)
            formats.write_wav16(p, pcm[u])
            paths.append(p)
        scp = os.path.join(d, "all.scp")
        open(scp, "w").write("\n".join(paths) + "\n")
        env = {"PKB_PRECISION": args.precision}
        clis = []
        if args.ref:
            clis.append(("reference CLI (CPU)", os.path.join(ROOT, "oracle/_ref/pocketkaldi_ref"), ()))
        clis.append(("shim CLI (GPU acoustic half, one utterance at a time)",
                     os.path.join(ROOT, "oracle/_ref/pocketkaldi_b200_cli"), ()))
        extra = ("--threads", str(args.threads)) if args.threads > 0 else ()
        clis.append(("batch CLI (one GPU batch, decoder on host threads)",
                     os.path.join(ROOT, "oracle/_ref/pocketkaldi_b200_batch"), extra))
        clis.append(("batch CLI --compact 0 (FP32 rows to the host decoder threads)",
                     os.path.join(ROOT, "oracle/_ref/pocketkaldi_b200_batch"), extra + ("--compact", "0")))
        clis.append(("batch CLI --gpu-decode 1 (acoustic half + Viterbi on the GPU)",
                     os.path.join(ROOT, "oracle/_ref/pocketkaldi_b200_batch"), extra + ("--gpu-decode", "1")))
        base = None
        audio = args.utts * args.seconds
        for name, cli, ex in clis:
            dt, hyps = run(cli, conf, scp, env, ex)
            if base is None:
                base = hyps
            agree = sum(a == b for a, b in zip(hyps, base))
            print("%-62s %8.2f s  RTFx %8.1f  hyps equal to first: %d/%d" %
                  (name, dt, audio / dt, agree, len(base)))


if __name__ == "__main__":
    main()
