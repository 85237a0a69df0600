#!/usr/bin/env python
"""GPU diagnostic for the tcgen05 conv kernel: dumps the raw accumulators of one pass of the first work
item and compares them with a float64 convolution of the same fp16 operands.  Test tooling (uses the
golden fixtures and the oracle); run on the B200 box:  python tools/diag_forward.py --pass 0 [--swap 1]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import _lib, batched as bk  # noqa: E402


def folded(sd):
    ws, bs = [], []
    for i in (0, 3, 6, 9, 12, 15, 18):
        w, b = torch.from_numpy(sd[f"conv.{i}.weight"]).double(), torch.from_numpy(sd[f"conv.{i}.bias"]).double()
        s = torch.from_numpy(sd[f"conv.{i+1}.weight"]).double() / torch.sqrt(torch.from_numpy(sd[f"conv.{i+1}.running_var"]).double() + 1e-5)
        ws.append((w * s[:, None, None, None]).float().half().double())
        bs.append(((b - torch.from_numpy(sd[f"conv.{i+1}.running_mean"]).double()) * s + torch.from_numpy(sd[f"conv.{i+1}.bias"]).double()).float().double())
    return ws, bs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pass", dest="ps", type=int, default=0)
    ap.add_argument("--swap", type=int, default=0)   # kept for old command lines; ignored
    ap.add_argument("--boards", type=int, default=10)
    a = ap.parse_args()
    g = os.path.join(ROOT, "tests", "golden")
    P, N, sd = dict(np.load(os.path.join(g, "positions.npz"))), dict(np.load(os.path.join(g, "nets.npz"))), dict(np.load(os.path.join(g, "weights_policy_17.npz")))
    dev = torch.device("cuda", 0)
    src = N["src"][: a.boards]
    pos = bk.Positions.from_numpy(P["board"][src], P["ko"][src], P["last"][src], P["turn"][src], dev)
    out = bk.features_batch(pos, want=("conv", "u8"))
    B = len(src)
    pol = bk.PackedNet(sd, dev)
    L = _lib.lib()
    dump = torch.full((640, 128), float("nan"), dtype=torch.float32, device=dev)
    logits = torch.zeros(B, 81, device=dev); probs = torch.zeros(B, 81, device=dev)
    flags = 1
    rc = L.bk_forward_debug(_lib.ptr(out["conv"]), _lib.ptr(pol.blob), None, _lib.ptr(logits), _lib.ptr(probs), None, B, flags,
                            _lib.stream_ptr(dev), _lib.ptr(dump), a.ps, None)
    print("launch rc", rc)
    try:
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        w = (C.c_uint * 8)()
        L.bk_debug_words(w)
        print("KERNEL FAILED:", e)
        print("debug words:", [hex(x) for x in w])
        return 1
    d = dump.cpu().double()
    ws, bs = folded(sd)
    x = torch.from_numpy(P["feats"][src[:5]]).double().reshape(-1, 27, 9, 9)
    acc = F.conv2d(x, ws[0], None, padding=2) + bs[0][None, :, None, None]   # layer-0 accumulators incl. the bias rows, group 0
    if a.ps in (0, 1):
        exp = torch.full((640, 128), float("nan"), dtype=torch.float64)
        for b in range(x.shape[0]):
            for p in range(81):
                exp[121 * b + 22 + 11 * (p // 9) + p % 9] = acc[b, :, p // 9, p % 9]
        rows = range(0, 384) if a.ps == 0 else range(384, 640)      # layer 0 runs as tiles 0..2, then tiles 3..4
    else:
        h = acc
        for l in range(1, a.ps):                                   # pass ps (>=2) is layer ps-1
            h = F.relu(h).float().half().double()
            h = F.conv2d(h, ws[l], None, padding=1) + bs[l][None, :, None, None]
        exp = torch.full((640, 128), float("nan"), dtype=torch.float64)
        for b in range(x.shape[0]):
            for p in range(81):
                exp[100 * b + 10 + 10 * (p // 9) + p % 9] = h[b, :, p // 9, p % 9]
        rows = range(0, 512)
    rows = [r for r in rows if not torch.isnan(exp[r, 0])]
    got, want = d[rows], exp[rows]
    err = (got - want).abs()
    print(f"pass {a.ps} swap {a.swap}: rows checked {len(rows)}, nan in dump {int(torch.isnan(got).sum())}, "
          f"max |acc| {float(want.abs().max()):.3f}, max err {float(err.max()):.4e}, rows ok {(err.max(1).values < 1e-2).sum().item()}")
    bad = torch.nonzero(err.max(1).values >= 1e-2).flatten()[:8].tolist()
    for i in bad:
        print("  bad row", rows[i], "got", got[i, :4].tolist(), "want", want[i, :4].tolist())
    wl = torch.from_numpy(N["logits17"][:B])
    print("final logits max err vs golden:", float((logits.cpu() - wl).abs().max()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
