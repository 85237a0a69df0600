#!/usr/bin/env python
"""GPU stress for the persistent playout kernel's hand-overs (step phase -> feature operand -> layer 0, activation-buffer
scratch, resident positions): the same batch of self-play / --simulate games played over and over, every other launch with the
L2 flushed, must give bit-identical move records, positions and liberty caches.  python tools/stress_playout.py [--iters 200]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import _lib, batched as bk, playout as po  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--batches", default="37,148,512,740,745,1485")
    a = ap.parse_args()
    g = os.path.join(ROOT, "tests", "golden")
    dev = torch.device("cuda", 0)
    p17 = bk.PackedNet(dict(np.load(os.path.join(g, "weights_policy_17.npz"))), dev)
    p19 = bk.PackedNet(dict(np.load(os.path.join(g, "weights_policy_19.npz"))), dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    L = _lib.lib()
    launches = 0
    for B in [int(x) for x in a.batches.split(",")]:
        for mode in (bk.MODE_SELFPLAY, bk.MODE_MCTS):
            ref = None
            try:
                for it in range(a.iters if B <= 745 else max(10, a.iters // 4)):
                    if it % 2 == 0:
                        flush.zero_()
                    pos = bk.Positions.empty(B, dev, track_libs=False)
                    res = po.run_playouts(pos, p17, mode, seed=9, game0=3, policy_odd=p19 if mode == bk.MODE_SELFPLAY else None, persistent=True)
                    launches += 1
                    got = (res.moves, pos.boards, pos.libs, pos.turn, res.score)
                    if ref is None:
                        torch.cuda.synchronize()
                        ref = [t.clone() for t in got]
                    elif not all(torch.equal(x, y) for x, y in zip(got, ref)):
                        bad = torch.nonzero((got[0] != ref[0]).any(1)).flatten()
                        raise AssertionError(f"iter {it}: {len(bad)} games differ: {bad[:12].tolist()}")
                torch.cuda.synchronize()
                print(f"B={B} mode={mode}: ok", flush=True)
            except Exception as e:  # noqa: BLE001
                w = (C.c_uint * 8)()
                L.bk_debug_words(w)
                print(f"B={B} mode={mode}: FAILED: {e}")
                print("debug words:", [hex(x) for x in w])
                return 1
    print(f"{launches} persistent launches (72 / 81 moves of 7-8 conv passes each), all bit-identical to the first of their kind")
    return 0


if __name__ == "__main__":
    sys.exit(main())
