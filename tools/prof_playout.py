#!/usr/bin/env python
"""Clock stamps of the persistent playout kernel (bk_playout_run_debug): for the first item of CTA 0, per move, where the time
between the last layer of one policy evaluation and the first layer of the next goes -- softmax, the move (table, legality
flags, sampling, play, new table, planes), the hand-over to the tensor pipe.  Prints cycles (SM clock)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import _lib, batched as bk  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    dev = torch.device("cuda", 0)
    g = os.path.join(ROOT, "tests", "golden")
    p17 = bk.PackedNet(dict(np.load(os.path.join(g, "weights_policy_17.npz"))), dev)
    p19 = bk.PackedNet(dict(np.load(os.path.join(g, "weights_policy_19.npz"))), dev)
    L = _lib.lib()
    n_steps = 72 if mode == 1 else 81
    for rep in range(2):
        pos = bk.Positions.empty(B, dev)
        moves = torch.empty(n_steps, B, dtype=torch.int16, device=dev)
        prof = torch.zeros(32 * 16, dtype=torch.int64, device=dev)
        rc = L.bk_playout_run_debug(_lib.ptr(pos.boards), _lib.ptr(pos.ko), _lib.ptr(pos.last), _lib.ptr(pos.turn), _lib.ptr(pos.libs),
                                    _lib.ptr(pos.done), _lib.ptr(p17.blob), _lib.ptr(p19.blob), C.c_uint64(1),
                                    C.c_uint32(0), mode, 70 if mode == 1 else 80, 0, n_steps, 1, _lib.ptr(moves), B, _lib.stream_ptr(dev),
                                    _lib.ptr(prof))
        assert rc == 0
        torch.cuda.synchronize()
    p = prof.cpu().numpy().reshape(32, 16)
    names = ["acc(L6)->epilogue done+sync", "softmax+sync", "table", "sampling", "play+publish", "new table", "planes",
             "zero+fence+sync+arrive", "arrive -> MMA warp resumes", "resume -> first pass accumulators", "whole move (L0 acc -> next L0 acc)"]
    rows = []
    for k in range(2, 30):
        s, n = p[k], p[k + 1]
        rows.append([s[6] - s[9], s[0] - s[6], s[1] - s[0], s[2] - s[1], s[3] - s[2], s[4] - s[3], s[5] - s[4], s[8] - s[5], n[11] - s[8],
                     n[10] - n[11], n[10] - s[10]])
    r = np.array(rows, np.float64)
    print(f"boards {B} mode {mode}: median / max cycles over moves 2..29 of CTA 0's first board")
    for i, nm in enumerate(names):
        print(f"  {nm:42s} {np.median(r[:, i]):9.0f} {r[:, i].max():9.0f}")
    step = r[:, :8].sum(1) + r[:, 8]
    print(f"  last-layer accumulators -> MMA warp resumes  {np.median(step):9.0f}  = {100 * np.median(step / r[:, 10]):.1f} % of a move")


if __name__ == "__main__":
    main()
