#!/usr/bin/env python
"""The split tail of the persistent 3x3 training kernel against the same kernel without it (BK_TC_NO_TAIL=1 is read once per process, so two
processes):   python tools/check_train_tail.py dump <file.npz> [P ...]   /   python tools/check_train_tail.py compare <a.npz> <b.npz>
Logits and gradients of a forward + backward per size; the comparison prints the largest difference per size relative to the largest
entry of each tensor.  Only the order of the four channel-group partial sums differs between the two, so they agree to round-off."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if sys.argv[1] == "dump":
    from bokego_b200 import reinforce as rf
    dev = torch.device("cuda", 0)
    g = os.path.join(ROOT, "tests", "golden")
    sd17 = dict(np.load(os.path.join(g, "weights_policy_17.npz")))
    G = np.load(os.path.join(g, "reinforce.npz"))
    calls = np.concatenate([G["black3/calls"], G["white2/calls"]])
    prec = int(os.environ.get("PREC", "5"))
    out = {}
    for P in [int(x) for x in sys.argv[3:]] or [190, 200, 220, 576]:
        planes = torch.from_numpy(np.ascontiguousarray(calls[np.arange(P) % len(calls)])).to(dev)
        rng = np.random.default_rng(P)
        moves = torch.from_numpy(rng.integers(0, 81, P)).to(dev).to(torch.int16)
        coef = torch.from_numpy(rng.uniform(-1, 1, P)).to(dev).to(torch.float32)
        tr = rf.PolicyTrainer(sd17, dev, prec=prec)
        logits = tr.forward(planes)[0]
        tr.backward(moves, coef)
        out[f"{P}/logits"] = logits.cpu().numpy()
        for k, v in tr.grads_dict().items():
            out[f"{P}/{k}"] = v.numpy()
    np.savez(sys.argv[2], **out)
    print("saved", sys.argv[2], len(out), "arrays")
else:
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    sizes = sorted({k.split("/")[0] for k in a.files}, key=int)
    for P in sizes:
        worst, where = 0.0, ""
        for k in a.files:
            if not k.startswith(P + "/"):
                continue
            scale = float(np.abs(b[k]).max())
            if scale < 1e-4:          # the conv biases in front of a BatchNorm: their gradient is zero up to round-off
                continue
            e = float(np.abs(a[k] - b[k]).max()) / scale
            if e > worst:
                worst, where = e, k
        print(f"P={P}: largest difference {worst:.3e} of the tensor's largest entry ({where})")
