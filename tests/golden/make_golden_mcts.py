#!/usr/bin/env python
"""Golden tree statistics of the REFERENCE search (bokego/mcts.py, unmodified, imported from /root/reference) driven by a
deterministic stand-in for the two nets (tests/fake_nets.py: a function of the position only).  With the nets taken out
of the picture the visit counts are an exact, reproducible function of the search rule, so the array-based search of
bokego_b200.mcts can be compared count for count.  Runs only in the build container:

    python tests/golden/make_golden_mcts.py        ->  tests/golden/mcts.npz
"""
import os
import sys

import numpy as np
import torch
from torch.distributions.categorical import Categorical

REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import bokego.go as go            # noqa: E402  (the reference)
import bokego.mcts as mcts        # noqa: E402

sys.path.insert(0, os.path.dirname(HERE))
from fake_nets import fake_nets   # noqa: E402  (tests/fake_nets.py: pure numpy helper, no device needed)

ENC = {go.BLACK: 1, go.WHITE: -1, go.EMPTY: 0}
DEC = {1: go.BLACK, -1: go.WHITE, 0: go.EMPTY}


def _arr(game):
    return np.array([[ENC[c] for c in game.board]], np.int8), np.array([game.turn], np.int16)


def fake_policy_dist(policy, game, device=None, fts=None):
    p, _ = fake_nets(*_arr(game))
    return Categorical(torch.from_numpy(p[0]))


def fake_value(v, game, device=None, fts=None):
    _, val = fake_nets(*_arr(game))
    return torch.tensor(val[0]).item()          # float32 -> Python float, like nnet.value's .item()


mcts.policy_dist = fake_policy_dist           # mcts.py:10 imported the name
mcts.nnet.value = fake_value                  # mcts.py:399 calls nnet.value


def run(board, ko, last, turn, n_rollouts, expand_thresh):
    for c in (mcts.MCTS._val_cache, mcts.MCTS._dist_cache, mcts.MCTS._fts_cache):
        c.clear()
    root = mcts.Go_MCTS(board="".join(DEC[int(v)] for v in board), ko=None, turn=int(turn),
                        last_move=None if last == -2 else int(last))
    dummy = torch.nn.Linear(1, 1)
    tree = mcts.MCTS(root, dummy, dummy, no_sim=True, expand_thresh=expand_thresh)
    tree.rollout(n_rollouts)
    visits, vsum = np.zeros(81, np.int64), np.zeros(81, np.float64)
    for ch in tree.children[tree.root]:
        visits[ch.last_move], vsum[ch.last_move] = tree.N[ch], tree.V[ch]
    n_nodes = sum(len(v) for v in tree.children.values()) + 1
    root_n, root_v = tree.N[tree.root], tree.V[tree.root]
    best = tree.choose()
    return visits, vsum, n_nodes, root_n, root_v, best.last_move, len(mcts.MCTS._val_cache)


def main():
    P = dict(np.load(os.path.join(HERE, "positions.npz")))
    # roots: the empty board and a few mid-game positions without ko (the reference cannot hash a fresh node with a ko)
    cand = [i for i in range(len(P["board"])) if P["ko"][i] < 0 and P["last"][i] >= 0 and 10 <= P["turn"][i] <= 60]
    rng = np.random.RandomState(7)
    roots = [None] + [int(i) for i in rng.choice(cand, 5, replace=False)]
    cases = []
    for r in roots:
        for n_roll, thresh in ((300, 100), (400, 3), (250, 0)):
            cases.append((r, n_roll, thresh))
    out = {k: [] for k in ("board", "ko", "last", "turn", "n_rollouts", "expand_thresh", "visits", "vsum", "n_nodes", "root_n",
                           "root_v", "best", "n_value_evals")}
    for r, n_roll, thresh in cases:
        if r is None:
            bd, ko, last, turn = np.zeros(81, np.int8), -1, -2, 0
        else:
            bd, ko, last, turn = P["board"][r], -1, int(P["last"][r]), int(P["turn"][r])
        v, vs, nn, rn, rv, best, ne = run(bd, ko, last, turn, n_roll, thresh)
        for k, x in zip(out, (bd, ko, last, turn, n_roll, thresh, v, vs, nn, rn, rv, best, ne)):
            out[k].append(x)
        print(f"root {r} rollouts {n_roll} thresh {thresh}: nodes {nn} evals {ne} best {best} top visits {np.sort(v)[-3:]}")
    np.savez_compressed(os.path.join(HERE, "mcts.npz"), **{k: np.array(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
