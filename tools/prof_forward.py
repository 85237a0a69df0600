#!/usr/bin/env python
"""GPU diagnostic: per-pass clock64 stamps of CTA 0 of the tcgen05 conv kernel (bk_forward_debug).
Prints, per pass, the MMA phase (issue start -> accumulators ready), the epilogue phase (accumulators ready ->
operands written) and the hand-over gaps, in SM cycles.  Run on the B200 box:
    python tools/prof_forward.py [--batch 4096] [--nets both|policy]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import _lib, batched as bk  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--nets", default="both")
    ap.add_argument("--flags", type=int, default=0)
    a = ap.parse_args()
    g = os.path.join(ROOT, "tests", "golden")
    P, sd = dict(np.load(os.path.join(g, "positions.npz"))), dict(np.load(os.path.join(g, "weights_policy_17.npz")))
    dev = torch.device("cuda", 0)
    n = len(P["board"])
    idx = np.arange(a.batch) % n
    pos = bk.Positions.from_numpy(P["board"][idx], P["ko"][idx], P["last"][idx], P["turn"][idx], dev)
    out = bk.features_batch(pos, want=("conv",))
    pol = bk.PackedNet(sd, dev)
    L = _lib.lib()
    B = a.batch
    logits = torch.zeros(B, 81, device=dev)
    probs = torch.zeros(B, 81, device=dev)
    value = torch.zeros(B, device=dev)
    prof = torch.zeros(1024, dtype=torch.int64, device=dev)
    flags = (1 if a.nets == "policy" else 3) | a.flags
    for rep in range(3):
        prof.zero_()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        rc = L.bk_forward_debug(_lib.ptr(out["conv"]), _lib.ptr(pol.blob), _lib.ptr(pol.blob) if flags & 2 else None,
                                _lib.ptr(logits), _lib.ptr(probs), _lib.ptr(value) if flags & 2 else None, B, flags,
                                _lib.stream_ptr(dev), None, -1, _lib.ptr(prof))
        ev1.record()
        torch.cuda.synchronize()
    print("rc", rc, "kernel ms", ev0.elapsed_time(ev1))
    pall = prof.cpu().numpy()
    p, w = pall[:256].reshape(64, 4), pall[256:512].reshape(64, 4)
    we, full, iss = pall[512:640], pall[640:768], pall[768:896]
    if iss[0]:
        print('peer CTA, first stages: [wait-slot start, copy issue, landed+forwarded] relative; copy latency')
        for i in range(0, 100, 1):
            print(f'  stage {i:3d}  wait_start {we[i]-iss[0]:8d}  issue {iss[i]-iss[0]:8d}  fwd {full[i]-iss[0]:8d}  latency {full[i]-iss[i]:6d}')
    used = [i for i in range(64) if p[i, 0] != 0]
    t0 = p[used[0], 0]
    print("pass  issue_start  issue_len  mma_phase(start->acc)  epilogue  gap_to_next_start | wait_own_w  wait_peer_w  issue")
    mma_tot = epi_tot = gap_tot = 0
    for k, i in enumerate(used):
        s, e, acc, act = p[i]
        nxt = p[used[k + 1], 0] if k + 1 < len(used) else act
        print(f"{i:4d} {s - t0:12d} {e - s:10d} {acc - s:12d} {act - acc:16d} {nxt - act:10d} | {w[i,0]:9d} {w[i,1]:9d} {w[i,2]:9d}")
        mma_tot += acc - s; epi_tot += act - acc; gap_tot += nxt - act
    tot = p[used[-1], 3] - t0
    print(f"passes {len(used)}  total {tot}  mma {mma_tot} ({100 * mma_tot / tot:.1f}%)  epilogue {epi_tot} "
          f"({100 * epi_tot / tot:.1f}%)  gaps {gap_tot} ({100 * gap_tot / tot:.1f}%)")


if __name__ == "__main__":
    main()
