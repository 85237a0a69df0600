#!/usr/bin/env python
"""debug: layer-0 / layer-1 conv outputs of the tcgen05 path against the FFMA path, read from the workspace"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import reinforce as rf
dev = torch.device("cuda", 0)
g = os.path.join(ROOT, "tests", "golden")
sd17 = dict(np.load(os.path.join(g, "weights_policy_17.npz")))
calls = np.load(os.path.join(g, "reinforce.npz"))["black3/calls"]
P = 4
planes = torch.from_numpy(np.ascontiguousarray(calls[40:40 + P])).to(dev)
def z_of(prec, layer):
    tr = rf.PolicyTrainer(sd17, dev, prec=prec)
    tr.forward(planes)
    torch.cuda.synchronize()
    ws = tr._ws.view(torch.float32)
    r32 = lambda n: (n + 31) // 32 * 32
    off = r32(P * 81 * 32)
    act = r32(P * 81 * 128)
    per = 2 * act + 2 * r32(P * 128)
    o = off + layer * per
    return ws[o:o + P * 81 * 128].reshape(P * 81, 128).cpu().numpy().copy()
for layer in (0, 1):
    ref = z_of(2, layer)
    for prec in (4, 5):
        z = z_of(prec, layer)
        print(f"layer {layer} prec {prec}: nan {np.isnan(z).sum()} zeros {(z == 0).mean():.3f} absmax {np.nanmax(np.abs(z)):.3f} ref absmax {np.abs(ref).max():.3f} maxerr {np.nanmax(np.abs(z - ref)):.3e}")
        if layer == 0 and prec == 4:
            np.set_printoptions(precision=3, suppress=True, linewidth=200)
            print("ref[0:3, 0:12]\n", ref[0:3, 0:12]); print("got[0:3, 0:12]\n", z[0:3, 0:12])
            print("ref[40, 0:12]\n", ref[40, 0:12]); print("got[40, 0:12]\n", z[40, 0:12])
            # is it a permutation / scaling?  correlate rows and columns
            rr = ref[:128]; zz = z[:128]
            cm = np.corrcoef(rr.T, zz.T)[:128, 128:]
            print("best column match for got col 0..7:", np.nanargmax(np.abs(cm), 0)[:8], np.nanmax(np.abs(cm), 0)[:8])
            rm = np.corrcoef(rr, zz)[:128, 128:]
            print("best row match for got row 0..7:", np.nanargmax(np.abs(rm), 0)[:8], np.nanmax(np.abs(rm), 0)[:8])
