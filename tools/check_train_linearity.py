#!/usr/bin/env python
"""Is the backward of the training kernels deterministic, and exactly linear in the coefficients?
    [BOKEGO_B200_SO=...] python tools/check_train_linearity.py [positions] [chunk] [prec]
The same 1,024 positions as tests/test_gpu_train.py::test_properties_at_scale: gradients twice with the same coefficients (must be
bit-identical), then with the coefficients doubled (every entry must double bit for bit: all the arithmetic is scale-invariant
under powers of two); per parameter tensor the number of entries that differ and the largest difference."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    prec = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    from bokego_b200 import reinforce as rf
    dev = torch.device("cuda", 0)
    g = os.path.join(ROOT, "tests", "golden")
    sd17 = dict(np.load(os.path.join(g, "weights_policy_17.npz")))
    G = np.load(os.path.join(g, "reinforce.npz"))
    calls = np.concatenate([G["black3/calls"], G["white2/calls"]])
    planes = torch.from_numpy(np.ascontiguousarray(calls[np.arange(P) % len(calls)])).to(dev)
    rng = np.random.default_rng(11)
    moves = torch.from_numpy(rng.integers(0, 81, P)).to(dev).to(torch.int16)
    coef = torch.from_numpy(rng.uniform(-1, 1, P)).to(dev).to(torch.float32)
    coef[rng.integers(0, P, 300)] = 0.0
    tr = rf.PolicyTrainer(sd17, dev, prec=prec)

    def grads(c):
        rf.compute_grads(tr, planes, moves, c, chunk=chunk)
        torch.cuda.synchronize()
        return {k: v.clone() for k, v in tr.grads_dict().items()}

    g1, g1b, g2 = grads(coef), grads(coef), grads(2 * coef)
    bad = 0
    for what, a, b, f in (("repeat", g1, g1b, 1.0), ("doubled", g1, g2, 2.0)):
        for k in a:
            d = (f * a[k] - b[k]).abs()
            n = int((d != 0).sum())
            if n:
                bad += 1
                print(f"{what:8s} {k:28s} {n:8d} of {d.numel():8d} entries differ, worst {float(d.max()):.3e} "
                      f"(largest entry {float(b[k].abs().max()):.3e})")
    print(f"P={P} chunk={chunk} prec={prec} lib={os.environ.get('BOKEGO_B200_SO', 'default')}: "
          + ("deterministic and exactly linear" if not bad else f"{bad} tensor comparisons differ"))


if __name__ == "__main__":
    main()
