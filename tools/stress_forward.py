#!/usr/bin/env python
"""GPU stress for the conv kernel's barrier protocol: many back-to-back launches at batch sizes that mix whole and split items;
on a trapped wait prints the debug words (thread, block, barrier, parity, tag).  python tools/stress_forward.py [--iters 300]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import _lib, batched as bk  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--batches", default="1,16,81,740,745,4096")
    ap.add_argument("--max-bad", type=int, default=6, help="stop after this many differing launches (measurement builds: count them all)")
    ap.add_argument("--quiet", action="store_true", help="one summary line per batch size")
    a = ap.parse_args()
    g = os.path.join(ROOT, "tests", "golden")
    P, sd = dict(np.load(os.path.join(g, "positions.npz"))), dict(np.load(os.path.join(g, "weights_policy_17.npz")))
    dev = torch.device("cuda", 0)
    pol = bk.PackedNet(sd, dev)
    L = _lib.lib()
    for B in [int(x) for x in a.batches.split(",")]:
        idx = np.arange(B) % len(P["board"])
        pos = bk.Positions.from_numpy(P["board"][idx], P["ko"][idx], P["last"][idx], P["turn"][idx], dev)
        conv = bk.features_batch(pos, want=("conv",))["conv"]
        ref = None
        n_bad = 0
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        try:
            for it in range(a.iters):
                if it % 2 == 0:
                    flush.zero_()          # cold L2: the weights of the first items come from HBM, which shifts all hand-overs
                for nets in ((pol, pol), (pol, None)):
                    l, p, v = bk.policy_value_batch(conv, B, nets[0], nets[1])
                    if ref is None:
                        torch.cuda.synchronize()
                        ref = l.clone()
                    elif not torch.equal(l, ref):
                        bad = torch.nonzero((l != ref).any(1)).flatten()
                        d = (l - ref).abs()
                        if not a.quiet:
                          print(f"B={B} iter {it} nets={'policy+value' if nets[1] is not None else 'policy'} cold={it % 2 == 0}: "
                              f"{len(bad)} rows differ {bad[:12].tolist()}, max abs diff {float(d.max()):.3e}, "
                              f"differing logits per row {(l != ref).sum(1)[bad[:6]].tolist()}, nan {int(torch.isnan(l).sum())}")
                        n_bad += 1
                        if n_bad >= a.max_bad:
                            raise AssertionError("result changed between launches")
            torch.cuda.synchronize()
            print(f"B={B}: {a.iters} x 2 launches " + ("ok" if n_bad == 0 else f"{n_bad} DIFFERENT"))
        except Exception as e:  # noqa: BLE001
            w = (C.c_uint * 8)()
            L.bk_debug_words(w)
            print(f"B={B}: FAILED: {e}")
            print("debug words:", [hex(x) for x in w])
            return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
