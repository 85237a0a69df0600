#!/usr/bin/env python
"""BASELINE configs[2]: MCTS genmove from the empty board, 1600 playouts per move, batched leaf evaluation under virtual loss.
Reports seconds per genmove and playouts/s for the reference's default search parameters (expand_thresh 100: the tree is
shallow and almost every rollout ends in an already evaluated leaf) and for an expand-on-second-visit search (expand_thresh 1:
every rollout reaches a new leaf, the regime where batching the leaves matters).  Run on the B200 box:
    python tools/bench_mcts.py > gpurun_out/mcts.jsonl"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import batched as bk, mcts  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    g = os.path.join(ROOT, "tests", "golden")
    sd17 = dict(np.load(os.path.join(g, "weights_policy_17.npz")))
    sdv = dict(np.load(os.path.join(g, "weights_policy_19.npz")))
    sdv.update(dict(np.load(os.path.join(g, "weights_value_head_standin.npz"))))   # seeded stand-in value head
    pol, val = bk.PackedNet(sd17, dev), bk.PackedNet(sdv, dev)
    mcts.MCTS(None, pol, val, device=dev).rollout(10)      # warm-up (kernel attributes, allocator)
    for thresh in (100, 1):
        for lb in (1, 8, 32, 128):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tree = mcts.MCTS(None, pol, val, expand_thresh=thresh, leaf_batch=lb, device=dev)
            tree.rollout(1600)
            mv = tree.choose()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(json.dumps({"config": "genmove, empty 9x9 board, 1600 playouts", "expand_thresh": thresh, "leaf_batch": lb,
                              "seconds": dt, "playouts_per_s": 1600 / dt, "move": mv, "nodes": int(tree.n),
                              "net_evals": tree.n_evals, "eval_batches": tree.n_eval_batches}))


if __name__ == "__main__":
    main()
