#!/usr/bin/env python
"""tcgen05 training GEMMs (prec 4 = TF32, 5 = 3xTF32) against the FFMA validation path on the same device: logits, every
gradient tensor, timing.  Run under `timeout`: a protocol bug ends in a trap after ~2 s, not in a hang."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import reinforce as rf  # noqa: E402

dev = torch.device("cuda", 0)
g = os.path.join(ROOT, "tests", "golden")
sd17 = dict(np.load(os.path.join(g, "weights_policy_17.npz")))
calls = np.load(os.path.join(g, "reinforce.npz"))["black3/calls"]
precs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [5, 4]
for P in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["45", "576"])]:
    planes = torch.from_numpy(np.ascontiguousarray(calls[np.arange(P) % len(calls)])).to(dev)
    rng = np.random.default_rng(3)
    moves = torch.from_numpy(rng.integers(0, 81, P).astype(np.int16)).to(dev)
    coef = torch.from_numpy(rng.uniform(-1, 1, P).astype(np.float32)).to(dev)
    ref = rf.PolicyTrainer(sd17, dev, prec=rf.PREC_FFMA)
    lref, _, _ = ref.forward(planes)
    ref.backward(moves, coef)
    torch.cuda.synchronize()
    gref = rf.tensors_from_flat(ref.grads)
    for prec in precs:
        tr = rf.PolicyTrainer(sd17, dev, prec=prec)
        l, _, _ = tr.forward(planes)
        torch.cuda.synchronize()
        print(f"P={P} prec={prec}: logits max err {float((l - lref).abs().max()):.3e}", flush=True)
        tr.backward(moves, coef)
        torch.cuda.synchronize()
        gm = rf.tensors_from_flat(tr.grads)
        worst = ("", 0.0)
        for k, v in gref.items():
            sc = float(np.abs(v).max())
            if sc < 1e-4:
                continue
            e = float(np.abs(gm[k] - v).max()) / sc
            if e > worst[1]:
                worst = (k, e)
        cos = float(torch.dot(tr.grads.double(), ref.grads.double()) / (tr.grads.double().norm() * ref.grads.double().norm()))
        print(f"P={P} prec={prec}: worst gradient error {worst[1]:.3e} ({worst[0]}), cosine {cos:.7f}", flush=True)
        for name, fn in (("forward", lambda: tr.forward(planes)), ("backward", lambda: tr.backward(moves, coef))):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                fn()
            b.record(); torch.cuda.synchronize()
            print(f"P={P} prec={prec}: {name} {a.elapsed_time(b) / 5:.3f} ms", flush=True)
