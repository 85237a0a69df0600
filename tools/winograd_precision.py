#!/usr/bin/env python
"""CPU study for the next round: would Winograd F(2x2,3x3) with fp16 operands / fp32 accumulation hold the parity bar?
Emulates the six 3x3 layers of PolicyNet(policy_17) with transformed weights U = G g G^T and transformed inputs V = B^T d B
both rounded to fp16, products accumulated in fp32, output transform in fp32, next-layer activations rounded to fp16 -- and
compares the logits with the fp32 reference and with the direct fp16-operand scheme the kernel implements today."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nets as onets  # noqa: E402

G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float64)
BT = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float64)
AT = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float64)


def h(x):
    return x.float().half().double()


def fold(sd):
    ws, bs = [], []
    for i in onets.CONV_IDX:
        w, b = torch.from_numpy(sd[f"conv.{i}.weight"]).double(), torch.from_numpy(sd[f"conv.{i}.bias"]).double()
        s = torch.from_numpy(sd[f"conv.{i+1}.weight"]).double() / torch.sqrt(torch.from_numpy(sd[f"conv.{i+1}.running_var"]).double() + 1e-5)
        ws.append(w * s[:, None, None, None])
        bs.append((b - torch.from_numpy(sd[f"conv.{i+1}.running_mean"]).double()) * s + torch.from_numpy(sd[f"conv.{i+1}.bias"]).double())
    return ws, bs


def wino_layer(x, w, round_v=True, round_u=True):
    """x [B,C,9,9] (fp16 values), w [O,C,3,3] -> [B,O,9,9] via F(2x2,3x3) on a 10x10 output grid (5x5 tiles)"""
    B, C = x.shape[:2]
    U = torch.einsum("ij,ocjk,lk->ocil", G, w, G)                       # [O,C,4,4]
    if round_u:
        U = h(U)
    xp = F.pad(x, (1, 2, 1, 2))                                            # 12x12: pad 1 + one extra row/col for the 10x10 grid
    tiles = xp.unfold(2, 4, 2).unfold(3, 4, 2)                             # [B,C,5,5,4,4]
    V = torch.einsum("ij,bcxyjk,lk->bcxyil", BT, tiles, BT)
    if round_v:
        V = h(V)
    M = torch.einsum("ocil,bcxyil->boxyil", U, V)                          # fp32-like accumulate (double here)
    Y = torch.einsum("ij,boxyjk,lk->boxyil", AT, M, AT)                     # [B,O,5,5,2,2]
    Y = Y.permute(0, 1, 2, 4, 3, 5).reshape(B, -1, 10, 10)[:, :, :9, :9]
    return Y


def run(sd, x, mode):
    ws, bs = fold(sd)
    a = h(x.double())
    a = F.conv2d(a, h(ws[0]), None, padding=2) + bs[0][None, :, None, None]
    a = h(F.relu(a))
    for l in range(1, 7):
        if mode == "direct":
            y = F.conv2d(a, h(ws[l]), None, padding=1)
        else:
            y = wino_layer(a, ws[l])
        y = F.relu(y + bs[l][None, :, None, None])
        a = h(y) if l < 6 else y
    hw = torch.from_numpy(sd["conv.21.weight"]).double().reshape(1, 128, 1, 1)
    hb = torch.from_numpy(sd["conv.21.bias"]).double().reshape(1, 81)
    return (a * hw).sum(1).reshape(-1, 81) + hb


def main():
    g = os.path.join(ROOT, "tests", "golden")
    sd = dict(np.load(os.path.join(g, "weights_policy_17.npz")))
    P, N = dict(np.load(os.path.join(g, "positions.npz"))), dict(np.load(os.path.join(g, "nets.npz")))
    rng = np.random.default_rng(0)
    idx = rng.choice(len(P["feats"]), 400, replace=False)
    x = onets.planes_to_float(P["feats"][idx])
    ref = onets.policy_logits(sd, x).double()
    for mode in ("direct", "winograd"):
        out = run(sd, x, mode)
        d = (out - ref).abs()
        pe = (torch.softmax(out, 1) - torch.softmax(ref, 1)).abs().max()
        agree = (out.argmax(1) == ref.argmax(1)).double().mean()
        print(f"{mode:9s}: logits max err {float(d.max()):.4f} mean {float(d.mean()):.5f}  probs max err {float(pe):.2e}  argmax agreement {float(agree):.4f}")


if __name__ == "__main__":
    main()
