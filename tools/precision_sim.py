"""CPU simulation of tensor-core precision policies for the nnet (no GPU needed).

Emulates `tcgen05.mma kind::f16` (16-bit operands, exact products, FP32 accumulate) with torch
float32 matmuls over operands rounded to FP16 / BF16, and compares the resulting scaled
log-likelihoods (src/am.cc:106-112) with a float64 evaluation of the same net. A policy is a
per-layer list of (activation planes, weight planes): (1,1) = one MMA per product, (1,2) =
a_hi*(w_hi+w_lo) and (2,1) = (a_hi+a_lo)*w_hi two MMAs, (2,2) = three MMAs (BF16X3 / FP16X3),
"c" = that operand's first-order correction through FP8 (E4M3) planes at half the cost (FP16C8).
This is the experiment behind the choice of PKB_PREC_FP16C8 as the default precision of bench.py
(DESIGN.md section 1).

Usage: python tools/precision_sim.py [--net 3|4] [--utts N] [--policy NAME ...]
Test tooling only: it uses the CPU oracle for the front end.
"""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pocketkaldi_b200 import formats  # noqa: E402
from pocketkaldi_b200.synth import synth_pcm, synth_global_cmvn  # noqa: E402


def rnd(x, fmt):
    if fmt == "f32":
        return x
    dt = torch.float16 if fmt == "fp16" else torch.bfloat16
    return x.to(dt).to(torch.float32)


def split(x, fmt, planes):
    hi = rnd(x, fmt)
    if planes == 1:
        return [hi]
    return [hi, rnd(x - hi, fmt)]


def forward(feats, layers, prior, policy, fmt):
    """feats: [T][440] float32 spliced; policy: list of (a_planes, w_planes) per linear layer."""
    x = feats
    li = 0
    for l in layers:
        if l[0] == "linear":
            ap, wp = policy[li]
            li += 1
            W = torch.from_numpy(l[1])
            b = torch.from_numpy(l[2])
            if fmt == "f64":
                x = x.double() @ W.double().T + b.double()
                continue
            if ap == "c" or wp == "c":
                # FP16 main product + first-order corrections through FP8 (e4m3) operands
                a16 = rnd(x, "fp16"); w16 = rnd(W, "fp16")
                f8 = lambda t, k: (t * 2.0 ** k).to(torch.float8_e4m3fn).to(torch.float32) * 2.0 ** -k
                y = a16 @ w16.T
                if ap == "c":
                    y = y + f8(x - a16, 11) @ f8(w16, 4).T
                if wp == "c":
                    y = y + f8(a16, 0) @ f8(W - w16, 15).T
                x = y + b
                continue
            A = split(x, fmt, ap)
            Ws = split(W, fmt, wp)
            y = A[0] @ Ws[0].T
            if ap == 2:
                y = y + A[1] @ Ws[0].T
            if wp == 2:
                y = y + A[0] @ Ws[1].T
            x = y + b
        elif l[0] == "relu":
            x = torch.clamp_min(x, 0)
        elif l[0] == "normalize":
            d = x.shape[1]
            x = x * torch.sqrt(d / (x * x).sum(1, keepdim=True))
        elif l[0] == "softmax":
            x = torch.log_softmax(x, 1)
    lp = torch.log(torch.from_numpy(prior).to(x.dtype))
    return torch.clamp_min(x, float(np.log(1e-20))) - lp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", type=int, default=3)
    ap.add_argument("--utts", type=int, default=8)
    ap.add_argument("--normalize", action="store_true")
    ap.add_argument("--prior", default="uniform")
    args = ap.parse_args()
    from oracle.oracle import Oracle
    orc = Oracle()
    rng = np.random.default_rng(0)
    if args.net == 3:
        H, W_, P = 6, 1024, 3000
    else:
        H, W_, P = 7, 2048, 8000
    layers = formats.make_dnn(rng, 440, W_, H, P, normalize=args.normalize)
    if args.prior == "uniform":
        prior = np.full(P, 1.0 / P, np.float32)
    else:
        prior = rng.uniform(0.5, 1.5, P).astype(np.float32)
        prior /= prior.sum()
    g = synth_global_cmvn()
    feats = []
    for u in range(args.utts):
        pcm = synth_pcm(1234, [u], 160000)[0]
        ft = orc.cmvn(orc.fbank(pcm.astype(np.float32)), g)
        feats.append(orc.splice(ft, 5, 5))
    X = torch.from_numpy(np.concatenate(feats, 0))
    n_lin = H + 1
    t0 = time.time()
    ref = forward(X, layers, prior, [(1, 1)] * n_lin, "f64")
    print("frames", X.shape[0], "ref f64 %.1fs" % (time.time() - t0))
    ref_arg = ref.argmax(1)
    top2 = torch.topk(ref, 2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])
    print("margin quantiles", [float(torch.quantile(margin, q)) for q in (0.001, 0.01, 0.1, 0.5)])

    macs = [440 * W_] + [W_ * W_] * (H - 1) + [W_ * P]
    tot = float(sum(macs))

    def cost(policy):
        u = lambda v: 0.5 if v == "c" else v - 1
        return sum(m * (1 + u(a) + u(w)) for m, (a, w) in zip(macs, policy)) / tot

    S, A2, W2, X3 = (1, 1), (2, 1), (1, 2), (2, 2)
    C = ("c", "c")
    pols = {
        "f32": ("f32", [S] * n_lin),
        "bf16": ("bf16", [S] * n_lin),
        "bf16x3": ("bf16", [X3] * n_lin),
        "fp16": ("fp16", [S] * n_lin),
        "fp16 last W2": ("fp16", [S] * H + [W2]),
        "fp16 last A2": ("fp16", [S] * H + [A2]),
        "fp16 last X3": ("fp16", [S] * H + [X3]),
        "fp16 all W2": ("fp16", [W2] * n_lin),
        "fp16 all A2": ("fp16", [A2] * n_lin),
        "fp16 last2 X3": ("fp16", [S] * (H - 1) + [X3, X3]),
        "fp16 first X3 last X3": ("fp16", [X3] + [S] * (H - 1) + [X3]),
        "fp16 hidden X3, last S": ("fp16", [X3] * H + [S]),
        "fp16x3": ("fp16", [X3] * n_lin),
        "fp16+c8 all": ("fp16", [C] * n_lin),
        "fp16c8 (first layer X3)": ("fp16", [X3] + [C] * (n_lin - 1)),
        "fp16+c8 hidden, last S": ("fp16", [C] * H + [S]),
        "fp16+c8 hid, last Wc": ("fp16", [C] * H + [(1, "c")]),
        "fp16+c8 hid, last Ac": ("fp16", [C] * H + [("c", 1)]),
        "hidden X3, last W2": ("fp16", [X3] * H + [W2]),
        "hidden X3, last A2": ("fp16", [X3] * H + [A2]),
    }
    for name, (fmt, pol) in pols.items():
        t0 = time.time()
        out = forward(X, layers, prior, pol, fmt).double()
        d = (out - ref).abs()
        agree = float((out.argmax(1) == ref_arg).double().mean())
        flips = int((out.argmax(1) != ref_arg).sum())
        print("%-24s cost %.2f  max|dLL| %.2e  mean %.2e  argmax %.5f (%d flips)  %.1fs" %
              (name, cost(pol), float(d.max()), float(d.mean()), agree, flips, time.time() - t0))


if __name__ == "__main__":
    main()
