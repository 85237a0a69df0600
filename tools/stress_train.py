#!/usr/bin/env python
"""Repeat the same forward + backward many times and require bit-identical logits and gradients every time: a race in the
producer / MMA / result-warp protocol of the tcgen05 training kernels would show up as a sporadic difference.
    python tools/stress_train.py [--iters 400]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import reinforce as rf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=400)
args = ap.parse_args()
dev = torch.device("cuda", 0)
g = os.path.join(ROOT, "tests", "golden")
sd17 = dict(np.load(os.path.join(g, "weights_policy_17.npz")))
calls = np.load(os.path.join(g, "reinforce.npz"))["black3/calls"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
bad = 0
for P in (16, 45, 576, 1100):
    planes = torch.from_numpy(np.ascontiguousarray(calls[np.arange(P) % len(calls)])).to(dev)
    rng = np.random.default_rng(P)
    moves = torch.from_numpy(rng.integers(0, 81, P).astype(np.int16)).to(dev)
    coef = torch.from_numpy(rng.uniform(-1, 1, P).astype(np.float32)).to(dev)
    for prec in (rf.PREC_TC_3XTF32, rf.PREC_TC_TF32):
        tr = rf.PolicyTrainer(sd17, dev, prec=prec)
        l0, _, _ = tr.forward(planes)
        tr.backward(moves, coef)
        g0, l0 = tr.grads.clone(), l0.clone()
        n_bad = 0
        for it in range(args.iters):
            if it % 7 == 0:
                flush.zero_()                      # cold L2 now and then
            l, _, _ = tr.forward(planes)
            tr.backward(moves, coef)
            if not (torch.equal(l, l0) and torch.equal(tr.grads, g0)):
                n_bad += 1
        torch.cuda.synchronize()
        print(f"P={P} prec={prec}: {args.iters} repeats, {n_bad} differ", flush=True)
        bad += n_bad
print("stress ok" if bad == 0 else f"STRESS FAILED: {bad}")
sys.exit(1 if bad else 0)
