import sys
sys.path.insert(0,'.')
import numpy as np
import pocketkaldi_b200 as pk
from oracle.oracle import Oracle
oracle=Oracle()
ctx=pk.Context(0)
rng = np.random.default_rng(23)
layers = []
d = 440
import os
for o in [int(v) for v in os.environ.get('WIDTHS','320,512,256').split(',')]:
    layers += [("linear", (rng.standard_normal((o, d)) * np.sqrt(2.0 / d)).astype(np.float32), (rng.standard_normal(o) * 0.1).astype(np.float32)), ("relu",), ("normalize",)]
    d = o
layers += [("linear", (rng.standard_normal((1000, d)) * np.sqrt(2.0 / d)).astype(np.float32), (rng.standard_normal(1000) * 0.1).astype(np.float32)), ("softmax",)]
prior = rng.uniform(0.5, 1.5, 1000).astype(np.float32); prior /= prior.sum()
feats = [(rng.standard_normal((n, 40)) * 2.5).astype(np.float32) for n in (300, 1, 420)]
ref = [oracle.am_compute(f, layers, prior, 5, 5) for f in feats]
import os
prec = getattr(pk, os.environ.get("PREC", "PREC_FP16C8"))
if os.environ.get("NONORM"):
    layers = [l for l in layers if l[0] != "normalize"]
    ref = [oracle.am_compute(f, layers, prior, 5, 5) for f in feats]
for it in range(5):
    am = pk.AcousticModel(ctx, prec).from_layers(layers, prior, 5, 5)
    outs=am.compute_batch(feats)
    outs2=am.compute_batch(feats)
    same=all(np.array_equal(a,b) for a,b in zip(outs,outs2))
    msg=["second call identical: %s" % same]
    for o, r in zip(outs, ref):
        bad=np.where(o.argmax(1)!=r.argmax(1))[0]
        err=np.abs(o-r).max(1)
        big=np.where(err>1e-2)[0]
        msg.append("bad %d rows %s maxerr %.2e" % (len(bad), (big.min(), big.max()) if len(big) else None, err.max()))
    print(it, " | ".join(msg))
    am.close()
