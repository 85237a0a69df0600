#!/usr/bin/env python
"""Two REINFORCE steps on 576 positions per precision (the workload ncu captures for the training kernels)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import reinforce as rf  # noqa: E402

dev = torch.device("cuda", 0)
g = os.path.join(ROOT, "tests", "golden")
sd17 = dict(np.load(os.path.join(g, "weights_policy_17.npz")))
calls = np.load(os.path.join(g, "reinforce.npz"))["black3/calls"]
P = int(sys.argv[1]) if len(sys.argv) > 1 else 576
planes = torch.from_numpy(np.ascontiguousarray(calls[np.arange(P) % len(calls)])).to(dev)
moves = torch.randint(0, 81, (P,), device=dev).to(torch.int16)
coef = torch.full((P,), 1.0 / 16, device=dev)
for prec in (rf.PREC_TC_3XTF32, rf.PREC_TC_TF32):
    tr = rf.PolicyTrainer(sd17, dev, prec=prec)
    for _ in range(2):
        rf.reinforce_step(tr, planes, moves, coef)
torch.cuda.synchronize()
print("ok")
