#!/usr/bin/env python
"""A short program for ncu: one self-play batch through the persistent playout kernel and one through the launch-per-move loop
(512 boards by default, policy_17 vs policy_19).  usage: python tools/run_playout_once.py [boards] [moves]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import batched as bk, playout as po  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
moves = int(sys.argv[2]) if len(sys.argv) > 2 else 72
dev = torch.device("cuda", 0)
g = os.path.join(ROOT, "tests", "golden")
p17 = bk.PackedNet(dict(np.load(os.path.join(g, "weights_policy_17.npz"))), dev)
p19 = bk.PackedNet(dict(np.load(os.path.join(g, "weights_policy_19.npz"))), dev)
for persistent in (True, False):
    pos = bk.Positions.empty(B, dev, track_libs=False)
    res = po.run_playouts(pos, p17, bk.MODE_SELFPLAY, seed=1, policy_odd=p19, n_steps=moves, persistent=persistent, graph=False)
    torch.cuda.synchronize()
    print("persistent" if persistent else "launch per move", "black wins", int((res.reward > 0).sum()), "of", B)
