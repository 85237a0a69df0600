#!/usr/bin/env python
"""Timing of the REINFORCE row on one B200: train-mode forward, backward, AdamW, and whole `reinforce` iterations.
    python tools/bench_train.py [--positions 576 2048] [--out gpurun_out/train.jsonl]
CUDA events on the launching stream, warm-up first.  Algorithmic work per position (valid taps):
forward 133.4 MFLOP, weight gradient 133.4 MFLOP, data gradient 122.9 MFLOP (no layer-0 data gradient)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FLOP_FWD = 2 * 66_706_944
FLOP_BWD = 2 * (66_706_944 + 6 * 10_240_000 + 10_368)


def timed(fn, n, warm=2):
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--positions", type=int, nargs="+", default=[36, 576, 2048])
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--precs", type=int, nargs="+", default=[5, 4, 0, 1, 2])
    ap.add_argument("--no-iterations", action="store_true", help="skip the whole reinforce() iterations")
    args = ap.parse_args()
    from bokego_b200 import reinforce as rf, batched as bk, nnet
    dev = torch.device("cuda", 0)
    g = os.path.join(ROOT, "tests", "golden")
    sd17 = dict(np.load(os.path.join(g, "weights_policy_17.npz")))
    sd19 = dict(np.load(os.path.join(g, "weights_policy_19.npz")))
    calls = np.load(os.path.join(g, "reinforce.npz"))["black3/calls"]
    lines = []
    for P in args.positions:
        planes = torch.from_numpy(np.ascontiguousarray(calls[np.arange(P) % len(calls)])).to(dev)
        moves = torch.randint(0, 81, (P,), device=dev).to(torch.int16)
        coef = torch.full((P,), 1.0 / 16, device=dev)
        for prec, name in ((5, "tc_3xtf32"), (4, "tc_tf32"), (0, "tf32"), (1, "3xtf32"), (2, "ffma")):
            if prec not in args.precs:
                continue
            tr = rf.PolicyTrainer(sd17, dev, prec=prec)
            f = timed(lambda: tr.forward(planes), args.iters)
            tr.forward(planes)
            b = timed(lambda: tr.backward(moves, coef), args.iters)
            s = timed(lambda: rf.reinforce_step(tr, planes, moves, coef), args.iters)
            line = {"positions": P, "prec": name, "forward_ms": f, "backward_ms": b, "step_ms": s,
                    "positions_per_s": P / (1e-3 * s), "forward_tflops": FLOP_FWD * P / (1e-3 * f) / 1e12,
                    "backward_tflops": FLOP_BWD * P / (1e-3 * b) / 1e12}
            print(json.dumps(line), flush=True)
            lines.append(line)
    # whole iterations of the reference's loop: bs games of self-play (pi in train mode) + the step
    pi, opp = nnet.PolicyNet(), nnet.PolicyNet()
    pi.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd17.items()})
    opp.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd19.items()})
    pi.to(dev).train()
    opp.to(dev).eval()
    for bs, acc in (() if args.no_iterations else ((16, "reference"), (16, "batch"), (256, "batch"))):
        opt = torch.optim.AdamW(pi.parameters(), lr=1e-5)
        rf.reinforce(pi, opp, opt, "black", n_itrs=1, bs=bs, device=dev, stats=[], accumulate=acc)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 5
        rf.reinforce(pi, opp, opt, "black", n_itrs=n, bs=bs, device=dev, stats=[], accumulate=acc)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        line = {"reinforce_iteration": {"bs": bs, "accumulate": acc, "seconds": dt, "games_per_s": bs / dt}}
        print(json.dumps(line), flush=True)
        lines.append(line)
    if args.out:
        with open(args.out, "w") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")


if __name__ == "__main__":
    main()
