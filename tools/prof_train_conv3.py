#!/usr/bin/env python
"""Clock stamps of the persistent 3x3 training kernel (CTA 0), from a measurement build:
    bash tools/build_variant.sh r3prof "-DBK_R3_PROF=1" bk_train_tc
    BOKEGO_B200_SO=$PWD/bokego_b200/libbokego_b200_r3prof.so python tools/prof_train_conv3.py [positions]
A train-mode forward; the stamps are those of its last 3x3 layer.  Per slab of the MMA-issuing warp: cycles spent waiting for the
group buffer, for the weight stage, for the accumulator, and issuing; per slab of the weight loader: waiting for the stage to be
handed back; per chain of the result warps: waiting for the chain, reading it out."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import reinforce as rf, _lib  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 576
dev = torch.device("cuda", 0)
g = os.path.join(ROOT, "tests", "golden")
sd17 = dict(np.load(os.path.join(g, "weights_policy_17.npz")))
calls = np.load(os.path.join(g, "reinforce.npz"))["black3/calls"]
planes = torch.from_numpy(np.ascontiguousarray(calls[np.arange(P) % len(calls)])).to(dev)
tr = rf.PolicyTrainer(sd17, dev, prec=rf.PREC_TC_3XTF32)
for _ in range(3):
    tr.forward(planes)
torch.cuda.synchronize()
L = _lib.lib()
if "--backward" in sys.argv:                      # also a backward pass: stamps of the MMA issuer of the last 3x3 weight gradient (CTA (0, 0))
    moves = torch.randint(0, 81, (P,), device=dev).to(torch.int16)
    tr.backward(moves, torch.full((P,), 1.0 / 16, device=dev))
    torch.cuda.synchronize()
buf = np.zeros(5 * 1024, dtype=np.int64)
L.bk_r3_prof_read.restype = C.c_int
L.bk_r3_prof_read.argtypes = [C.c_void_p]
assert L.bk_r3_prof_read(buf.ctypes.data) == 0
mma, ld, res, wg, wp = (buf[i * 1024:(i + 1) * 1024].reshape(256, 4) for i in range(5))
nw = int((wg[:, 0] > 0).sum())
if nw:
    per = np.diff(wg[:nw, 0])
    print(f"weight gradient, MMA issuer of CTA (0, 0): {nw} slabs, period mean {per.mean():.0f} (median {np.median(per):.0f}); waits for the "
          f"producers' slab mean {(wg[:nw, 1] - wg[:nw, 0]).mean():.0f}, issue (incl. waiting for an accumulator) {(wg[:nw, 2] - wg[:nw, 1]).mean():.0f}")
    npd = int((wp[:, 0] > 0).sum())
    if npd > 1:
        print(f"weight gradient, producer warp 4: period mean {np.diff(wp[:npd, 0]).mean():.0f}; issuing the next slab's loads {(wp[:npd, 1] - wp[:npd, 0]).mean():.0f}, "
              f"waiting for the stage {(wp[:npd, 2] - wp[:npd, 1]).mean():.0f}, transposing / splitting / storing (incl. waiting for the loads) "
              f"{(wp[:npd, 3] - wp[:npd, 2]).mean():.0f}, rest {(np.diff(wp[:npd, 0]) - (wp[:npd - 1, 3] - wp[:npd - 1, 0])).mean():.0f}")
    print("  slab: period wait_full issue")
    for i in range(min(nw - 1, 46)):
        print(f"  {i:3d} {per[i]:6d} {wg[i, 1] - wg[i, 0]:6d} {wg[i, 2] - wg[i, 1]:6d}")
n = int((mma[:, 0] > 0).sum())
t0 = mma[0, 0]
print(f"P={P}: {n} slabs on CTA 0, {mma[n - 1, 3] - t0} cycles from the first slab's top to the last slab's issue")
w_a, w_w, w_c = mma[:n, 1] - mma[:n, 0], mma[:n, 2] - mma[:n, 1], mma[:n, 3] - mma[:n, 2]
period = np.diff(mma[:n, 0])
print(f"MMA warp per slab: period mean {period.mean():.0f} (median {np.median(period):.0f}); waits: group buffer {w_a.mean():.0f}, "
      f"weights {w_w.mean():.0f}, accumulator {w_c.mean():.0f}; issue + rest {(period.mean() - (w_a + w_w + w_c)[:-1].mean()):.0f}")
print("slab: top(rel) wait_A wait_W wait_acc")
for i in range(min(n, 80)):
    print(f"  {i:3d} {mma[i, 0] - t0:8d} {w_a[i]:6d} {w_w[i]:6d} {w_c[i]:6d}")
m = int((ld[:, 0] > 0).sum())
print(f"weight loader: {m} slabs, wait for the stage mean {(ld[:m, 1] - ld[:m, 0]).mean():.0f} cycles; period {np.diff(ld[:m, 0]).mean():.0f}")
for it in range(4):
    rows = res[it * 14:(it + 1) * 14]
    k = int((rows[:, 0] > 0).sum())
    if k == 0:
        break
    print(f"result warps, item {it}: " + " ".join(f"[wait {rows[c, 1] - rows[c, 0]} read {rows[c, 2] - rows[c, 1]}]" for c in range(k))
          + f" acc complete at {rows[13, 3] - t0}; first 32 columns in the buffer +{rows[13, 0] - rows[13, 3]}, written +{rows[13, 1] - rows[13, 3]},"
          f" all stores issued +{rows[13, 2] - rows[13, 3]}; first chain wait of the item starts at {rows[0, 0] - t0}")
