"""Whole playouts on the device (bokego_b200.playout): CUDA-graph replay vs step-by-step launches, the oracle
driven by the device's own probabilities, and invariance under sharding of the global game ids."""
import numpy as np
import pytest
import torch

from oracle import cpu as ocpu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(sd17, sd19):
    from bokego_b200 import batched as bk, playout as po
    dev = torch.device("cuda", 0)
    return bk, po, dev, bk.PackedNet(sd17, dev), bk.PackedNet(sd19, dev)


@pytest.mark.parametrize("mode", [0, 1])
def test_graph_replay_equals_stepwise(env, mode):
    bk, po, dev, p17, p19 = env
    B = 257
    odd = p19 if mode == 1 else None
    a = po.run_playouts(bk.Positions.empty(B, dev, track_libs=False), p17, mode, seed=11, game0=5, policy_odd=odd, graph=True)
    b = po.run_playouts(bk.Positions.empty(B, dev, track_libs=False), p17, mode, seed=11, game0=5, policy_odd=odd, graph=False)
    torch.cuda.synchronize()
    assert torch.equal(a.records(), b.records())
    rec = a.records().cpu().numpy()
    if mode == 1:
        assert (rec[:, 0] == 72).all()                   # selfplay.py:21-33: always 72 moves
        assert (rec[:, 3:] >= 0).all()
    else:
        assert ((rec[:, 0] > 80) | (rec[:, 3:] == -1).any(1)).all()   # terminal: turn > 80 or a PASS
    assert set(np.unique(rec[:, 1])) <= {-1, 1}


def test_selfplay_games_vs_oracle(env):
    """each step: the oracle gets the probabilities the device computed and must make the same move; at the end the
    boards, turn counts and results agree (bit-exact playouts given identical probabilities and draws)"""
    bk, po, dev, p17, p19 = env
    B, seed, game0 = 96, 77, 300
    pos = bk.Positions.empty(B, dev)
    bd = np.zeros((B, 81), np.int8); ko = np.full(B, -1, np.int16); last = np.full(B, -2, np.int16)
    turn = np.zeros(B, np.int16); done = np.zeros(B, np.uint8); libs = np.zeros((B, 81), np.uint8)
    bufs = {}
    for k in range(po.n_steps_for(bk.MODE_SELFPLAY, 70)):
        out = bk.features_batch(pos, fresh_libs=(k == 0), want=("conv", "libs"), out=bufs)
        _, probs, _ = bk.policy_value_batch(out["conv"], B, p17 if k % 2 == 0 else p19, None, want_logits=False)
        _, _, lo = ocpu.features_batch(bd, ko, last, turn, None if k == 0 else libs)
        libs[:] = lo
        mv = bk.playout_step(pos, probs, bk.MODE_SELFPLAY, 70, seed=seed, game0=game0).cpu().numpy()
        mo = ocpu.step_batch(bd, ko, last, turn, libs, done, probs.cpu().numpy(), 1, 70, seed=seed, game0=game0)
        assert np.array_equal(mv, mo), k
    assert done.all() and np.array_equal(pos.boards.cpu().numpy(), bd)
    res = po.run_playouts(bk.Positions.empty(B, dev, track_libs=False), p17, bk.MODE_SELFPLAY, seed=seed, game0=game0,
                          policy_odd=p19)
    assert np.array_equal(res.score.cpu().numpy().astype(np.float64), ocpu.score_batch(bd))
    assert np.array_equal(res.n_moves.cpu().numpy(), turn)


def test_results_do_not_depend_on_sharding(env):
    bk, po, dev, p17, p19 = env
    n = 203
    _, _, whole = po.self_play(n, p17, p19, dev, seed=5)
    parts = []
    for world in (2, 8):
        recs = [po.self_play(n, p17, p19, dev, seed=5, rank=r, world=world)[2].records() for r in range(world)]
        parts.append(torch.cat(recs))
    torch.cuda.synchronize()
    for p in parts:
        assert torch.equal(p, whole.records())
    lo, hi, sim = po.simulate(64, p17, dev, seed=9, rank=1, world=2)
    assert (lo, hi) == (32, 64) and sim.moves.shape[0] == 32
