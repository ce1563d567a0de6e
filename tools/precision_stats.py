"""Error distribution of the single-MMA precision modes against the split-BF16 parity mode
on the config-3 net (16 x 10 s). Usage on a B200: python tools/precision_stats.py"""
import numpy as np, sys
sys.path.insert(0,'.')
import pocketkaldi_b200 as pk
from pocketkaldi_b200 import formats
from pocketkaldi_b200.synth import synth_global_cmvn
ctx=pk.Context(0)
rng=np.random.default_rng(0)
layers=formats.make_dnn(rng,440,1024,6,3000)
prior=rng.uniform(0.5,1.5,3000).astype(np.float32); prior/=prior.sum()
g=synth_global_cmvn()
outs={}
for name,prec in (("bf16x3",pk.PREC_BF16X3),("fp16",pk.PREC_FP16),("bf16",pk.PREC_BF16)):
    am=pk.AcousticModel(ctx,prec).from_layers(layers,prior,5,5)
    b=pk.Batch(ctx,[160000]*48,g,am,prob_scale=0.1)
    b.synth_pcm(1234,0); b.run(pk.STAGE_ALL)
    outs[name]=b.get(pk.BUF_LOGLIK).reshape(48,998,3000); b.close(); am.close()
ref=outs["bf16x3"]
for n in ("fp16","bf16"):
    d=np.abs(outs[n]-ref)/0.1
    i=np.unravel_index(d.argmax(), d.shape); print("   argmax of error at (utt,frame,pdf)", i, "values", outs[n][i], ref[i])
    print(n,"max",d.max(),"q99.99",np.quantile(d,0.9999),"q99.9",np.quantile(d,0.999),"q99",np.quantile(d,0.99),"mean",d.mean(),"argmax",np.mean(outs[n].argmax(2)==ref.argmax(2)))
    # per-frame max
    fm=d.max(axis=2)
    print("   per-frame max: median",np.median(fm),"q99",np.quantile(fm,0.99),"frac frames > 2e-2",np.mean(fm>2e-2))
    # margin filtered argmax
    top2=np.sort(ref,axis=2)[:,:,-2:]; margin=top2[:,:,1]-top2[:,:,0]
    ok=margin>2e-2
    print("   argmax on frames with margin>2e-2:",np.mean((outs[n].argmax(2)==ref.argmax(2))[ok]),"frac frames",ok.mean())
