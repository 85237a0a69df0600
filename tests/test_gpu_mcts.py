"""Tree search over the hot path (bokego_b200.mcts, BASELINE configs[2]):
  * bk_make_moves (Go_MCTS.make_move) against the oracle's play_move for every square of random positions;
  * the array-based search against tree statistics of the REFERENCE search (tests/golden/mcts.npz, produced by
    tests/golden/make_golden_mcts.py with a deterministic stand-in for the nets): visit counts, value sums, node counts
    and the chosen move must agree exactly;
  * batched leaf evaluation under virtual loss with the real nets: bookkeeping invariants and agreement of the chosen move
    with the sequential search."""
import os

import numpy as np
import pytest
import torch

from fake_nets import fake_net_tree
from oracle import cpu as ocpu

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEC = {1: "X", -1: "O", 0: "."}


@pytest.fixture(scope="module")
def env(sd17, sd_value):
    from bokego_b200 import batched as bk, go, mcts
    dev = torch.device("cuda", 0)
    return bk, go, mcts, dev, bk.PackedNet(sd17, dev), bk.PackedNet(sd_value, dev)


def test_make_moves_vs_oracle(env, positions):
    bk, go, mcts, dev, _, _ = env
    rng = np.random.default_rng(3)
    pick = rng.choice(len(positions["board"]), 40, replace=False)
    bd, ko, last, turn = (positions[k][pick] for k in ("board", "ko", "last", "turn"))
    _, _, libs = ocpu.features_batch(bd, ko, last, turn)                  # exact liberties = the cache of a fresh Game
    pos = bk.Positions.from_numpy(bd, ko, last, turn, dev, libs)
    par = torch.arange(len(pick), dtype=torch.int32, device=dev).repeat_interleave(82)
    mv = torch.arange(-1, 81, dtype=torch.int16, device=dev).repeat(len(pick))
    child, status = bk.make_moves(pos, par, mv)
    cb, ck, cl, ct, clb, st = (t.cpu().numpy() for t in (child.boards, child.ko, child.last, child.turn, child.libs, status))
    for i in range(len(pick)):
        for m in range(-1, 81):
            c = i * 82 + m + 1
            s, b2, k2, l2, t2, lb2 = ocpu.play(bd[i], ko[i], last[i], turn[i], m, libs=libs[i])
            assert st[c] == s, (i, m)
            if s == 0:
                assert np.array_equal(cb[c], b2) and ck[c] == k2 and cl[c] == l2 and ct[c] == t2, (i, m)
                assert np.array_equal(clb[c], lb2), (i, m)
            else:
                assert np.array_equal(cb[c], bd[i]) and ct[c] == turn[i]


def test_sequential_search_equals_reference_tree(env):
    bk, go, mcts, dev, _, _ = env
    g = dict(np.load(os.path.join(GOLDEN, "mcts.npz")))
    for i in range(len(g["turn"])):
        board = "".join(DEC[int(v)] for v in g["board"][i])
        last = None if g["last"][i] == -2 else int(g["last"][i])
        root = go.Game(board=board, ko=None, last_move=last, turn=int(g["turn"][i]))
        tree = fake_net_tree(mcts)(root, expand_thresh=int(g["expand_thresh"][i]), device=dev)
        tree.rollout(int(g["n_rollouts"][i]))
        assert np.array_equal(tree.root_visits(), g["visits"][i]), i
        lo, c = tree.child0[tree.root], tree.nchild[tree.root]
        vs = np.zeros(81); vs[tree.move[lo: lo + c]] = tree.V[lo: lo + c]
        assert np.allclose(vs, g["vsum"][i], rtol=0, atol=1e-9), i
        assert tree.N[tree.root] == g["root_n"][i] and abs(tree.V[tree.root] - g["root_v"][i]) < 1e-9
        assert tree.n == g["n_nodes"][i], (i, tree.n, g["n_nodes"][i])
        assert tree.best_move() == g["best"][i]
        assert tree.choose() == g["best"][i] and tree.nchild[tree.root] >= 0


def test_batched_search_with_virtual_loss(env):
    bk, go, mcts, dev, pol, val = env
    seq = mcts.MCTS(None, pol, val, expand_thresh=2, leaf_batch=1, device=dev)
    seq.rollout(800)
    bat = mcts.MCTS(None, pol, val, expand_thresh=2, leaf_batch=32, device=dev)
    bat.rollout(1600)
    for t, n in ((seq, 800), (bat, 1600)):
        assert t.N[t.root] == n and t.root_visits().sum() == n      # the root is expanded from the start: every rollout enters a child
        assert np.isfinite(t.V[: t.n]).all() and (t.N[: t.n] >= 0).all()
        # every node's visits = visits of its children + the rollouts that stopped there
        for i in np.flatnonzero(t.nchild[: t.n] > 0)[:50]:
            lo, c = t.child0[i], t.nchild[i]
            assert t.N[lo: lo + c].sum() <= t.N[i]
    assert bat.n_eval_batches < bat.n_evals            # leaves really were evaluated in batches
    top = np.argsort(seq.root_visits())[::-1][:3]
    assert bat.best_move() in top
    assert 0.0 < bat.winrate() < 1.0
    mv = bat.choose()
    assert 0 <= mv < 81 and bat.nchild[bat.root] > 0


def test_simulate_search_equals_reference_tree(env):
    """--simulate mode (MCTS(no_sim=False), mcts.py:133-151, 195-217) against tree statistics of the unmodified reference
    search (tests/golden/mcts_sim.npz, make_golden_mcts_sim.py: fake nets, keyed draws): per root child N, Q and V, the root's
    own sums, node count, winrate and chosen move; playouts run through bk_playout_step on the device"""
    from fake_nets import fake_net_sim_tree
    bk, go, mcts, dev, _, _ = env
    g = dict(np.load(os.path.join(GOLDEN, "mcts_sim.npz")))
    Tree = fake_net_sim_tree(mcts, int(g["seed"]))
    for i in range(len(g["turn"])):
        board = "".join(DEC[int(v)] for v in g["board"][i])
        last = None if g["last"][i] == -2 else int(g["last"][i])
        root = go.Game(board=board, ko=None, last_move=last, turn=int(g["turn"][i]))
        tree = Tree(root, None, None if g["with_value"][i] else False, no_sim=False, expand_thresh=int(g["expand_thresh"][i]),
                    value_net_weight=float(g["w"][i]), device=dev)
        tree.rollout(int(g["n_rollouts"][i]))
        lo, c = tree.child0[tree.root], tree.nchild[tree.root]
        got = np.zeros((3, 81))
        got[0, tree.move[lo: lo + c]], got[1, tree.move[lo: lo + c]], got[2, tree.move[lo: lo + c]] = \
            tree.N[lo: lo + c], tree.Q[lo: lo + c], tree.V[lo: lo + c]
        assert np.array_equal(got[:2], g["stats"][i][:2]), i
        assert np.allclose(got[2], g["stats"][i][2], rtol=0, atol=1e-9), i
        rs = g["root_stats"][i]
        assert tree.N[tree.root] == rs[0] and tree.Q[tree.root] == rs[1] and abs(tree.V[tree.root] - rs[2]) < 1e-9, i
        assert tree.n == g["n_nodes"][i], (i, tree.n, g["n_nodes"][i])
        assert abs(tree.winrate() - g["winrate"][i]) < 1e-12, i
        # most visited child; the reference breaks ties in the order of a Python set, the batched tree by the lowest move
        assert got[0, tree.best_move()] == got[0].max() == g["stats"][i][0, int(g["best"][i])], i


def test_simulate_search_on_the_device(env):
    """--simulate with the real nets: leaf batches are played out together by run_playouts; bookkeeping invariants, and a
    search is reproducible for a given seed"""
    bk, go, mcts, dev, pol, val = env
    runs = []
    for _ in range(2):
        t = mcts.MCTS(None, pol, val, no_sim=False, expand_thresh=4, leaf_batch=16, seed=5, device=dev)
        t.rollout(160)
        assert t.N[t.root] == 160 and t.n_playouts == 160 and abs(t.Q[t.root]) <= 160
        assert t.value_net_weight == 0.5 and 0.0 < t.winrate() < 1.0
        for i in np.flatnonzero(t.nchild[: t.n] > 0)[:40]:
            lo, c = t.child0[i], t.nchild[i]
            assert t.N[lo: lo + c].sum() <= t.N[i] and abs(t.Q[i]) <= t.N[i]
        runs.append((t.root_visits().copy(), t.Q[: t.n].copy()))
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    # policy only (no value net): w = 0, selection on playout results alone
    t = mcts.MCTS(None, pol, None, no_sim=False, expand_thresh=4, leaf_batch=8, device=dev)
    t.rollout(64)
    assert t.value_net_weight == 0.0 and t.N[t.root] == 64 and not t.V[: t.n].any()
