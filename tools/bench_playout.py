#!/usr/bin/env python
"""Per-move cost of the device playout loop at the board counts of BASELINE configs[4] (4,096 self-play games sharded over
1 / 2 / 4 / 8 GPUs = 4,096 ... 512 boards per GPU): whole games through the CUDA graph, and the two kernels of a move timed
alone (warm L2, as inside the loop).  One JSON line per board count."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import batched as bk, playout as po  # noqa: E402


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = torch.device("cuda", 0)
    g = os.path.join(ROOT, "tests", "golden")
    p17 = bk.PackedNet(dict(np.load(os.path.join(g, "weights_policy_17.npz"))), dev)
    p19 = bk.PackedNet(dict(np.load(os.path.join(g, "weights_policy_19.npz"))), dev)
    sizes = [int(x) for x in sys.argv[1:]] or [512, 1024, 2048, 4096]
    for B in sizes:
        sp = po.PlayoutGraph(B, dev, p17, bk.MODE_SELFPLAY, seed=1, policy_odd=p19, persistent=False)
        ms = timed(sp.replay, n=5, warm=2)
        spk = po.PlayoutGraph(B, dev, p17, bk.MODE_SELFPLAY, seed=1, policy_odd=p19, persistent=True)
        ms_k = timed(spk.replay, n=5, warm=2)
        # a mid-game position set for the per-kernel timings
        pos = bk.Positions.empty(B, dev)
        bufs = bk.features_batch(pos, fresh_libs=True, want=("conv", "libs"), out={"libs": pos.libs})
        probs = torch.empty(B, 81, device=dev)
        for k in range(30):
            bk.policy_value_batch(bufs["conv"], B, p17, None, want_logits=False, probs_out=probs)
            bk.playout_step(pos, probs, bk.MODE_SELFPLAY, 70, seed=3, encode_into=bufs["conv"])
        keep = [t.clone() for t in (pos.boards, pos.ko, pos.last, pos.turn, pos.libs, pos.done)]

        def restore():
            for t, s in zip((pos.boards, pos.ko, pos.last, pos.turn, pos.libs, pos.done), keep):
                t.copy_(s)
        fwd = timed(lambda: bk.policy_value_batch(bufs["conv"], B, p17, None, want_logits=False, probs_out=probs), n=50)
        mv = torch.empty(B, dtype=torch.int16, device=dev)

        def step_enc():
            restore()
            bk.playout_step(pos, probs, bk.MODE_SELFPLAY, 70, seed=3, moves_out=mv, encode_into=bufs["conv"])

        def step_only():
            restore()
            bk.playout_step(pos, probs, bk.MODE_SELFPLAY, 70, seed=3, moves_out=mv)
        t_restore = timed(restore, n=50)
        t_se = timed(step_enc, n=50) - t_restore
        t_s = timed(step_only, n=50) - t_restore
        t_e = timed(lambda: bk.features_batch(pos, fresh_libs=False, want=("conv", "libs"), out=bufs), n=50)
        print(json.dumps({"boards": B, "persistent_kernel": {"selfplay_ms_per_batch": ms_k, "us_per_move": 1e3 * ms_k / 72,
                                                             "games_per_s": B / (1e-3 * ms_k)},
                          "selfplay_ms_per_batch": ms, "us_per_move": 1e3 * ms / 72, "games_per_s": B / (1e-3 * ms),
                          "launches_per_move": 2, "forward_policy_us": 1e3 * fwd, "step_encode_us": 1e3 * t_se,
                          "step_only_us": 1e3 * t_s, "encode_only_us": 1e3 * t_e}), flush=True)


if __name__ == "__main__":
    main()
