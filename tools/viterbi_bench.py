"""Times the GPU Viterbi (pkb_batch_decode) apart from the acoustic stages on the decode-demo graph.
Usage on a B200: python tools/viterbi_bench.py [--utts 128] [--words 40]"""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import pocketkaldi_b200 as pk  # noqa: E402
from pocketkaldi_b200 import formats  # noqa: E402
from pocketkaldi_b200.synth import synth_global_cmvn  # noqa: E402
from decode_demo import word_loop_graph  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=128)
ap.add_argument("--words", type=int, default=40)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
ctx = pk.Context(0)
rng = np.random.default_rng(0)
layers = formats.make_dnn(rng, 440, 1024, 6, 3000)
prior = np.full(3000, 1.0 / 3000, np.float32)
graph, tid2pdf = word_loop_graph(args.words, 3, 3000)
am = pk.AcousticModel(ctx, pk.PREC_FP16C8).from_layers(layers, prior, 5, 5, tid2pdf=np.asarray(tid2pdf, np.int32))
fst = pk.Fst(ctx, graph=graph)
b = pk.Batch(ctx, [160000] * args.utts, synth_global_cmvn(), am, prob_scale=0.1)
b.synth_pcm(1234, 0)
for rep in range(args.reps):
    ctx.sync()
    t0 = time.perf_counter()
    b.run(pk.STAGE_ALL | pk.STAGE_NO_FEATS)
    ctx.sync()
    t1 = time.perf_counter()
    hyps, wts = b.decode(fst)
    t2 = time.perf_counter()
    print("utts %d: acoustic %.1f ms, viterbi %.1f ms (%.1f us per frame step), words/utt %.1f, failed %d" % (
        args.utts, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t1) * 1e6 / 998,
        np.mean([len(h) for h in hyps if h is not None]), sum(h is None for h in hyps)))
