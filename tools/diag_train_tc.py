#!/usr/bin/env python
"""Error profile (max / 90th percentile / median per tensor) of the tcgen05 and mma.sync 3xTF32 gradients against the FFMA path on the
recorded reference iteration (black, 3 games): the instrument that showed long TMEM accumulation chains losing 1e-3 in the early layers."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import reinforce as rf
dev = torch.device("cuda", 0)
g = os.path.join(ROOT, "tests", "golden")
sd17 = dict(np.load(os.path.join(g, "weights_policy_17.npz")))
G = dict(np.load(os.path.join(g, "reinforce.npz")))
tag = "black3"
color, bs = int(G[f"{tag}/color"]), int(G[f"{tag}/bs"])
calls, rfrom = G[f"{tag}/calls"], int(G[f"{tag}/replay_from"])
lengths, results, moves = G[f"{tag}/lengths"], G[f"{tag}/results"], G[f"{tag}/moves"]
pos = [(i, j) for i, n in enumerate(lengths) for j in range(1 if color else 0, int(n), 2)]   # the training colour's turns, game by game
mv = np.array([moves[g_, j] for g_, j in pos], np.int16)
last = len(lengths) - 1                      # the reference differentiates the last game only (selfplay.py:86)
coef = np.array([(-results[last] if color else results[last]) / bs if g_ == last else 0.0 for g_, _ in pos], np.float32)
planes = torch.from_numpy(np.ascontiguousarray(calls[rfrom:])).to(dev)
mvd, cfd = torch.from_numpy(mv).to(dev), torch.from_numpy(coef).to(dev)
def grads(prec, sub=None):
    tr = rf.PolicyTrainer(sd17, dev, prec=prec)
    p, m, c = (planes, mvd, cfd) if sub is None else (planes[sub].contiguous(), mvd[sub].contiguous(), cfd[sub].contiguous())
    tr.forward(p); tr.backward(m, c); torch.cuda.synchronize()
    return tr.grads.clone(), tr
ref, _ = grads(2)
for prec in (5, 5, 1):
    gq, tr = grads(prec)
    d = rf.tensors_from_flat(gq - ref); r = rf.tensors_from_flat(ref)
    out = []
    for k in ("conv.0.weight", "conv.1.weight", "conv.1.bias", "conv.3.weight", "conv.4.weight", "conv.10.weight", "conv.18.weight", "conv.19.weight", "conv.21.weight"):
        am = np.abs(r[k]).max(); e = np.abs(d[k]).ravel() / am
        out.append(f"{k}: max {e.max():.1e} q90 {np.quantile(e, .9):.1e} q50 {np.quantile(e, .5):.1e}")
    print(f"prec {prec} vs FFMA (108 positions):\n   " + "\n   ".join(out), flush=True)
# only the positions with a coefficient
sub = torch.from_numpy(np.nonzero(coef)[0]).to(dev)
ref2, _ = grads(2, sub)
g5, _ = grads(5, sub)
d = rf.tensors_from_flat(g5 - ref2); r = rf.tensors_from_flat(ref2)
for k in ("conv.1.weight", "conv.10.weight"):
    am = np.abs(r[k]).max(); e = np.abs(d[k]).ravel() / am
    print(f"36 positions prec 5 vs FFMA {k}: max {e.max():.1e} q90 {np.quantile(e, .9):.1e}")
print("full vs subset FFMA:", float((ref - ref2).abs().max() / ref.abs().max()))
