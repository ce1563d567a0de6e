"""How large must the refinement margin of PKB_PREC_FP16R be?  (GPU tool)

A frame's best pdf can only change under the FP16 pass if the error DIFFERENCE between two of its
leading pdfs exceeds their distance. This measures, on the random-init nets of BASELINE.json, the
spread delta = max_i e_i - min_i e_i of the FP16-pass error e = ll_fp16 - ll_fp16x3 over the pdfs
within 0.1 of the frame's best one, next to the plain maximum error. Any observed margin below
2 * delta could hide a flip, so the margin has to be >= 2 * max delta.

Usage: python tools/refine_margin_stats.py [--net 3|4] [--utts N]
"""
import argparse
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import pocketkaldi_b200 as pk  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--net", default="3")
ap.add_argument("--utts", type=int, default=24)
args = ap.parse_args()
cfg = bench.CONFIGS[args.net]
ctx = pk.Context(0)
from pocketkaldi_b200.synth import synth_global_cmvn  # noqa: E402
g = synth_global_cmvn()
layers, prior = bench.make_layers(cfg), bench.uniform_prior(cfg)


def run(prec, margin=None):
    am = pk.AcousticModel(ctx, prec).from_layers(layers, prior, 5, 5)
    if margin is not None:
        am.set_refine_margin(margin)
    b = pk.Batch(ctx, [bench.SAMPLES_10S] * args.utts, g, am, prob_scale=1.0)
    b.synth_pcm(4321, 0)
    b.run(pk.STAGE_ALL)
    out = b.get(pk.BUF_LOGLIK)
    st = b.refine_stats()
    b.close()
    am.close()
    return out, st


ref, _ = run(pk.PREC_FP16X3)
fast, _ = run(pk.PREC_FP16R, 0.0)   # margin 0: the FP16 pass alone (exact ties aside)
e = fast - ref
top = ref.max(axis=1, keepdims=True)
lead = ref >= top - 0.1
spread = np.where(lead, e, -np.inf).max(axis=1) - np.where(lead, e, np.inf).min(axis=1)
part = np.partition(ref, ref.shape[1] - 2, axis=1)
margin = part[:, -1] - part[:, -2]
flips = fast.argmax(1) != ref.argmax(1)
print("net %s: %d frames, %d pdfs" % (args.net, ref.shape[0], ref.shape[1]))
print("FP16 pass: max |dLL| %.3e, flips %d (largest margin of a flipped frame %.3e)"
      % (np.abs(e).max(), flips.sum(), margin[flips].max() if flips.any() else 0.0))
print("error spread over the leading pdfs: max %.3e, p99.9 %.3e, p99 %.3e, median %.3e"
      % (spread.max(), np.quantile(spread, 0.999), np.quantile(spread, 0.99), np.median(spread)))
for m in (0.01, 0.02, 0.03, 0.04, 0.06):
    out, (rows, n) = run(pk.PREC_FP16R, m)
    fl = out.argmax(1) != ref.argmax(1)
    print("margin %.2f: %5.2f %% of frames recomputed, flips vs FP16X3 %d, max |dLL| %.3e"
          % (m, 100.0 * n / ref.shape[0], fl.sum(), np.abs(out - ref).max()))
