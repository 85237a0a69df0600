#!/usr/bin/env python
"""Duration of bk_forward (policy + value, and policy only) at batch sizes around whole rounds of the grid: 3,700 boards = 10
full rounds of the 74 CTA pairs, 4,070 = 11 full rounds, 4,096 = 11 rounds + 12 left-over items (the tail).  L2 flushed before
every launch, CUDA events."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import batched as bk  # noqa: E402

dev = torch.device("cuda", 0)
g = os.path.join(ROOT, "tests", "golden")
P = dict(np.load(os.path.join(g, "positions.npz")))
pol = bk.PackedNet(dict(np.load(os.path.join(g, "weights_policy_17.npz"))), dev)
val = bk.PackedNet(dict(np.load(os.path.join(g, "weights_policy_19.npz"))), dev, is_value=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sizes = [int(x) for x in sys.argv[1:]] or [1, 16, 37, 74, 148, 370, 740, 3700, 4070, 4096, 4440]
for B in sizes:
    idx = np.arange(B) % len(P["board"])
    pos = bk.Positions.from_numpy(P["board"][idx], P["ko"][idx], P["last"][idx], P["turn"][idx], dev)
    conv = bk.features_batch(pos, want=("conv",))["conv"]
    out = {"boards": B}
    for name, nets in (("policy+value", (pol, val)), ("policy", (pol, None))):
        v = torch.empty(B, device=dev) if nets[1] is not None else None
        for cold in (True, False):
            ts = []
            for it in range(13):
                if cold:
                    flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                bk.policy_value_batch(conv, B, nets[0], nets[1], want_logits=False)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            out[f"{name} {'cold' if cold else 'warm'} L2 us"] = round(1e3 * float(np.median(ts[3:])), 1)
    print(json.dumps(out), flush=True)
