#!/usr/bin/env python
"""Golden tree statistics of the REFERENCE search in --simulate mode (MCTS(no_sim=False): bokego/mcts.py:133-151, 195-217,
319-364, unmodified, imported from /root/reference).  Runs only in the build container:

    python tests/golden/make_golden_mcts_sim.py        ->  tests/golden/mcts_sim.npz

What is taken out of the picture, so that the statistics are an exact function of the search rule:
  * the nets: tests/fake_nets.py (a deterministic function of the position), patched in for policy_dist / nnet.value as in
    make_golden_mcts.py;
  * the random numbers: Categorical.sample is replaced by torch's own single-sample algorithm argmax(p / q) (SURVEY 8c shim 2)
    with q = the counter-based Exp(1) stream of the kernels, keyed (seed, playout number, turn of the position, try 0).  The SAME
    q serves every redraw of one move, so the move is argmax over the acceptable squares of p / q whatever the state of the
    reference's mutable distribution cache (mcts.py:357 zeroes rejected moves in the CACHED distribution, so a second playout
    through the same position would otherwise consume different draws than the first).  `_simulate`, `find_random_child`,
    `get_move`, `make_move`, `reward`, `_backpropagate`, `_puct_select` run as they are;
  * gnugo: absent, so reward() falls through to Game.score() (SURVEY 8c shim 4).
A case is dropped when the cache mutation reached the search itself (a zeroed prior of a legal own-eye child inside the
tree, mcts.py:226): that side effect is deliberately not replicated (bokego_b200/mcts.py docstring).
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
sys.path.insert(0, os.path.dirname(HERE))

import bokego.go as go            # noqa: E402  (the reference)
import bokego.mcts as mcts        # noqa: E402

from fake_nets import fake_nets   # noqa: E402
from oracle import cpu as ocpu    # noqa: E402

ENC = {go.BLACK: 1, go.WHITE: -1, go.EMPTY: 0}
DEC = {1: go.BLACK, -1: go.WHITE, 0: go.EMPTY}
SEED = 4242


class State:
    playout = -1          # number of the playout that is running
    rewards = []


def _arr(game):
    return np.array([[ENC[c] for c in game.board]], np.int8), np.array([game.turn], np.int16)


def fake_policy_dist(policy, game, device=None, fts=None):
    p, _ = fake_nets(*_arr(game))
    d = Categorical(torch.from_numpy(p[0]))
    d._turn = int(game.turn)
    return d


def fake_value(v, game, device=None, fts=None):
    _, val = fake_nets(*_arr(game))
    return torch.tensor(val[0]).item()


def keyed_sample(dist, sample_shape=torch.Size()):
    q = torch.from_numpy(ocpu.exp_draws(SEED, State.playout, dist._turn, 0))
    return torch.argmax(dist.probs / q)


_simulate = mcts.MCTS._simulate


def counted_simulate(self, node, gnu=False):
    State.playout += 1
    r = _simulate(self, node, gnu)
    State.rewards.append(r)
    return r


mcts.policy_dist = fake_policy_dist
mcts.nnet.value = fake_value
mcts.MCTS._simulate = counted_simulate          # a counter around the unmodified method
Categorical.sample = keyed_sample


def run(board, last, turn, n_rollouts, expand_thresh, w, with_value):
    for c in (mcts.MCTS._val_cache, mcts.MCTS._dist_cache, mcts.MCTS._fts_cache):
        c.clear()
    State.playout, State.rewards = -1, []
    root = mcts.Go_MCTS(board="".join(DEC[int(v)] for v in board), ko=None, turn=int(turn), last_move=None if last == -2 else int(last))
    dummy = torch.nn.Linear(1, 1)
    tree = mcts.MCTS(root, dummy, dummy if with_value else None, no_sim=False, expand_thresh=expand_thresh, value_net_weight=w)
    tree.rollout(n_rollouts)
    assert abs(tree.value_net_weight - (w if with_value else 0.0)) < 1e-12
    tainted = any(float(n.dist.probs[c.last_move]) == 0.0 for n, kids in tree.children.items() for c in kids)
    stats = np.zeros((3, 81), np.float64)
    for ch in tree.children[tree.root]:
        stats[:, ch.last_move] = tree.N[ch], tree.Q[ch], tree.V[ch]
    n_nodes = sum(len(v) for v in tree.children.values()) + 1
    root_stats = (tree.N[tree.root], tree.Q[tree.root], tree.V[tree.root])
    wr = tree.winrate()
    best = tree.choose()
    return stats, n_nodes, root_stats, wr, best.last_move, np.array(State.rewards, np.int8), tainted


def main():
    P = dict(np.load(os.path.join(HERE, "positions.npz")))
    cand = [i for i in range(len(P["board"])) if P["ko"][i] < 0 and P["last"][i] >= 0 and 10 <= P["turn"][i] <= 50]
    rng = np.random.RandomState(11)
    roots = [None] + [int(i) for i in rng.choice(cand, 4, replace=False)]
    keys = ("board", "last", "turn", "n_rollouts", "expand_thresh", "w", "with_value", "stats", "n_nodes", "root_stats", "winrate", "best",
            "rewards")
    out = {k: [] for k in keys}
    for r in roots:
        for n_roll, thresh, w, with_value in ((120, 100, 0.5, 1), (150, 2, 0.5, 1), (100, 0, 0.25, 1), (100, 3, 0.5, 0)):
            if r is None:
                bd, last, turn = np.zeros(81, np.int8), -2, 0
            else:
                bd, last, turn = P["board"][r], int(P["last"][r]), int(P["turn"][r])
            stats, nn, rs, wr, best, rew, tainted = run(bd, last, turn, n_roll, thresh, w, with_value)
            print(f"root {r} turn {turn} rollouts {n_roll} thresh {thresh} w {w} value {with_value}: nodes {nn} root N/Q/V {rs} "
                  f"winrate {wr:.3f} best {best} black wins {int((rew > 0).sum())}/{len(rew)}" + ("  [dropped: cache mutation reached the tree]" if tainted else ""))
            if tainted:
                continue
            pad = np.zeros(160, np.int8); pad[: len(rew)] = rew
            for k, x in zip(keys, (bd, last, turn, n_roll, thresh, w, with_value, stats, nn, np.array(rs, np.float64), wr, best, pad)):
                out[k].append(x)
    np.savez_compressed(os.path.join(HERE, "mcts_sim.npz"), seed=SEED, **{k: np.array(v) for k, v in out.items()})
    print("cases kept:", len(out["turn"]))


if __name__ == "__main__":
    main()
