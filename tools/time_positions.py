#!/usr/bin/env python
"""encode + forward as two launches against bk_forward_positions (one launch, planes computed on chip), L2 flushed before every
step, CUDA events around the step."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import _lib, batched as bk  # noqa: E402

dev = torch.device("cuda", 0)
g = os.path.join(ROOT, "tests", "golden")
P = dict(np.load(os.path.join(g, "positions.npz")))
pol = bk.PackedNet(dict(np.load(os.path.join(g, "weights_policy_17.npz"))), dev)
val = bk.PackedNet(dict(np.load(os.path.join(g, "weights_policy_19.npz"))), dev, is_value=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = _lib.lib()
for B in [int(x) for x in sys.argv[1:]] or [16, 256, 1024, 4096, 16384]:
    idx = np.arange(B) % len(P["board"])
    pos = bk.Positions.from_numpy(P["board"][idx], P["ko"][idx], P["last"][idx], P["turn"][idx], dev)
    feats = {"conv": torch.empty(L.bk_feats_conv_bytes(B), dtype=torch.uint8, device=dev), "legal": torch.empty(B, 81, dtype=torch.uint8, device=dev)}
    probs, value = torch.empty(B, 81, device=dev), torch.empty(B, device=dev)
    outs = {"legal": feats["legal"]}

    def two():
        bk.features_batch(pos, fresh_libs=True, want=("conv", "legal"), out=feats)
        bk.policy_value_batch(feats["conv"], B, pol, val, want_logits=False, probs_out=probs, value_out=value)

    def one():
        bk.evaluate_positions(pos, pol, val, fresh_libs=True, want=("legal",), out=outs, probs_out=probs, value_out=value)
    res = {"boards": B}
    ts = {"encode + forward (2 launches) us": [], "forward_positions (1 launch) us": []}
    for it in range(24):                       # interleaved, so that clock / power drift hits both alike
        for name, fn in (("encode + forward (2 launches) us", two), ("forward_positions (1 launch) us", one)):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts[name].append(a.elapsed_time(b))
    for name in ts:
        res[name] = round(1e3 * float(np.median(ts[name][4:])), 1)
    print(json.dumps(res), flush=True)
