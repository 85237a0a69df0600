"""Row (b') of the coverage contract: the reference's OWN callers -- bokego/mcts.py, bokego/gtp.py, bin/selfplay.py, boke.py,
unmodified, from baseline/_ref (tools/install_reference.sh) -- drive the B200 path through the mirror modules.

  * the reference's MCTS / GTP search over bokego_b200.nnet / bokego_b200.go on the GPU gives the same tree as
    bokego_b200.mcts (leaf_batch = 1), with and without batched expansion through the class-level caches (row a18);
  * Go_MCTS.find_random_child (the --simulate playout loop, mcts.py:195-206,319-364) and selfplay.playout / legal_sample
    (bin/selfplay.py:18-47) produce, draw for draw, the moves of the batched device stepping kernel;
  * boke.py itself answers a GTP session.
Skipped when no reference install is present (the GPU box only has what travelled in baseline/_ref).
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from bokego_b200 import _lib, batched as bk, dropin, go, nnet

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(dropin.find_reference() is None, reason="no reference install in baseline/_ref")]
DEV = torch.device("cuda", 0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LUT = {go.BLACK: 1, go.WHITE: -1, go.EMPTY: 0}


@pytest.fixture(scope="module")
def ref():
    mcts, gtp, selfplay = dropin.reference_callers()
    yield mcts, gtp, selfplay
    dropin.batch_expansions(mcts.MCTS, False)
    dropin.uninstall()


@pytest.fixture(scope="module")
def nets(sd17, sd19, sd_value):
    t = lambda sd: {k: torch.from_numpy(np.asarray(a)) for k, a in sd.items()}
    pi, pi2, v = nnet.PolicyNet(), nnet.PolicyNet(), nnet.ValueNet()
    pi.load_state_dict(t(sd17)); pi2.load_state_dict(t(sd19)); v.load_state_dict(t(sd_value))
    return pi.eval().to(DEV), pi2.eval().to(DEV), v.eval().to(DEV)


def _clear(mcts):
    for c in (mcts.MCTS._val_cache, mcts.MCTS._dist_cache, mcts.MCTS._fts_cache):
        c.clear()


def _root_visits(tree):
    out = np.zeros(81, np.int64)
    for c in tree.children[tree.root]:
        out[c.last_move] = tree.N[c]
    return out


def _ref_search(mcts, pi, v, n, root=None, **kw):
    _clear(mcts)
    n0 = _lib.launch_count
    tree = mcts.MCTS(root if root is not None else mcts.Go_MCTS(), pi, v, no_sim=True, device=DEV, **kw)
    tree.rollout(n)
    return tree, _lib.launch_count - n0


def test_reference_search_matches_batched_tree(ref, nets):
    """mcts.py:133-234 unchanged over the mirror == bokego_b200.mcts (leaf_batch 1): same visit counts at the root and below"""
    mcts = ref[0]
    pi, _, v = nets
    from bokego_b200 import mcts as bmcts
    for n_roll, kw in ((200, {}), (300, {"expand_thresh": 20})):
        tree, launches = _ref_search(mcts, pi, v, n_roll, **kw)
        ours = bmcts.MCTS(None, pi, v, device=DEV, leaf_batch=1, **kw)
        ours.rollout(n_roll)
        assert np.array_equal(_root_visits(tree), ours.root_visits()), (n_roll, kw)
        # the same positions went through the value net (the batched tree also scores its root, which the reference never does)
        assert len(mcts.MCTS._val_cache) == ours.n_evals - 1
        assert launches > 0


def test_prefill_caches_with_real_tree(ref, nets):
    """row a18: nnet.prefill_caches against the reference's real Go_MCTS / MCTS -- same tree, fewer device calls"""
    mcts = ref[0]
    pi, _, v = nets
    kw = {"expand_thresh": 20}
    plain, launches_plain = _ref_search(mcts, pi, v, 300, **kw)
    visits_plain = _root_visits(plain)
    vals_plain = dict(mcts.MCTS._val_cache)
    dists_plain = {k: d.probs.clone() for k, d in mcts.MCTS._dist_cache.items()}
    dropin.batch_expansions(mcts.MCTS)
    try:
        batched, launches_batched = _ref_search(mcts, pi, v, 300, **kw)
    finally:
        dropin.batch_expansions(mcts.MCTS, False)
    assert np.array_equal(_root_visits(batched), visits_plain)
    assert {n: batched.N[n] for n in batched.N} == {n: plain.N[n] for n in plain.N}
    assert launches_batched < launches_plain, (launches_batched, launches_plain)
    # what the caches hold is what the one-position path computed: values bit for bit, distributions to the last ulp
    for node, val in vals_plain.items():
        assert mcts.MCTS._val_cache[node] == val
    worst = max(float((mcts.MCTS._dist_cache[n].probs - p).abs().max()) for n, p in dists_plain.items())
    assert worst <= 2e-7, worst
    for node in vals_plain:
        assert torch.equal(mcts.MCTS._fts_cache[node], nnet.features(go.Game(node.board, node.ko, node.last_move, node.turn))) or \
            node._libs is not None


def test_gtp_genmove_200_rollouts(ref, nets):
    """BASELINE configs[0] through the reference's code path, nets on the GPU: GTP(time_lim=0, n_rollouts=200).send("genmove b")
    (gtp.py:199-214, 344-366)"""
    mcts, gtp_mod, _ = ref
    pi, _, v = nets
    _clear(mcts)
    g = gtp_mod.GTP(mcts.Go_MCTS(), pi, v, no_sim=True, time_lim=0, n_rollouts=200, pondering=False, device=DEV)
    g.running = True
    out = g.send("genmove b")
    assert out.startswith("= ") and out.endswith("\n\n")
    mv = go.squash(out[2:].strip())
    assert go.Game().is_legal(mv)
    assert g.root.turn == 1 and g.root.last_move == mv
    out = g.send("genmove w")
    assert out.startswith("= ")
    assert "=" in g.send("showboard")
    # the same move as the batched engine with one leaf per evaluation
    from bokego_b200 import mcts as bmcts
    ours = bmcts.MCTS(None, pi, v, device=DEV, leaf_batch=1)
    ours.rollout(200)
    assert ours.best_move() == mv


class _RecordingSampler:
    """torch's single-sample multinomial is argmax(p / q), q ~ Exp(1) (SURVEY 8c shim 2); this stand-in for
    Categorical.sample draws q from a seeded CPU generator and keeps it, so that the device kernel can be fed the same
    draws.  It is the sampling shim of tests/golden/make_golden.py, applied at run time."""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)
        self.draws = []

    def __call__(self, dist, sample_shape=torch.Size()):
        q = torch.empty(81, dtype=torch.float32).exponential_(1.0, generator=self.gen)
        self.draws.append(q)
        p = dist.probs
        return torch.argmax(p / q.to(p.device)).reshape(())


def _device_positions(game):
    bd = np.array([[LUT[c] for c in game.board]], np.int8)
    last = -1 if game.last_move == go.PASS else (-2 if not isinstance(game.last_move, int) else int(game.last_move))
    return bk.Positions.from_numpy(bd, [-1 if game.ko is None else int(game.ko)], [last], [int(game.turn)], DEV)


def test_find_random_child_playout_equals_device_stepping(ref, nets, monkeypatch):
    """the --simulate loop of the reference (mcts.py:195-206) over the mirror, against bk_playout_step fed the same draws"""
    mcts = ref[0]
    pi, _, v = nets
    from torch.distributions.categorical import Categorical
    for seed in (3, 11):
        _clear(mcts)
        rec = _RecordingSampler(seed)
        monkeypatch.setattr(Categorical, "sample", lambda d, s=torch.Size(), r=rec: r(d, s))
        tree = mcts.MCTS(mcts.Go_MCTS(), pi, v, no_sim=True, device=DEV)
        node, moves, used = tree.root, [], []
        while not node._terminal:
            n_before = len(rec.draws)
            nxt = node.find_random_child()
            moves.append(nxt.last_move)
            used.append(len(rec.draws) - n_before)
            node = nxt
        monkeypatch.undo()
        assert len(moves) >= 40
        # device side: one board, policy probabilities from the kernel, the recorded draws injected move by move
        pos = _device_positions(go.Game())
        pnet = pi._packed(DEV)
        at, dev_moves = 0, []
        for k, n_used in enumerate(used):
            out = bk.features_batch(pos, want=("conv", "libs"))
            _, probs, _ = bk.policy_value_batch(out["conv"], 1, pnet, None, want_logits=False)
            q = torch.stack(rec.draws[at: at + n_used] + [torch.ones(81)] * (82 - n_used)).reshape(1, 82, 81).to(DEV)
            mv = bk.playout_step(pos, probs, bk.MODE_MCTS, 80, q_inj=q.contiguous())
            dev_moves.append(int(mv[0]))
            at += n_used
        assert dev_moves == moves, seed
        assert "".join({1: go.BLACK, -1: go.WHITE, 0: go.EMPTY}[int(x)] for x in pos.boards[0].cpu()) == node.board
        r = node.reward()
        _, reward = bk.score_batch(pos.boards)
        assert int(reward[0]) == r


def test_selfplay_playout_equals_device_stepping(ref, nets, monkeypatch):
    """bin/selfplay.py:18-47 (playout / legal_sample) unchanged over the mirror, against the device self-play stepping"""
    selfplay = ref[2]
    if selfplay is None:
        pytest.skip("bin/selfplay.py not in the reference install")
    pi, pi2, _ = nets
    from torch.distributions.categorical import Categorical
    rec = _RecordingSampler(5)
    monkeypatch.setattr(Categorical, "sample", lambda d, s=torch.Size(), r=rec: r(d, s))
    g = go.Game(moves=[])
    selfplay.playout(g, pi, pi2, device=DEV)
    monkeypatch.undo()
    assert len(g.moves) == 72 and g.turn == 72
    pos = _device_positions(go.Game())
    p1, p2 = pi._packed(DEV), pi2._packed(DEV)
    dev_moves = []
    for k in range(72):
        out = bk.features_batch(pos, want=("conv", "libs"))
        _, probs, _ = bk.policy_value_batch(out["conv"], 1, p1 if k % 2 == 0 else p2, None, want_logits=False)
        q = rec.draws[k].reshape(1, 1, 81).to(DEV).contiguous()
        dev_moves.append(int(bk.playout_step(pos, probs, bk.MODE_SELFPLAY, 70, q_inj=q)[0]))
    assert dev_moves == list(g.moves)
    assert len(rec.draws) == 72


def test_boke_py_session(tmp_path, sd17, sd_value):
    """the reference's boke.py, unmodified, as a GTP engine process on the GPU (`-g`), nets from torch checkpoints"""
    root = dropin.find_reference()
    if not os.path.isfile(os.path.join(root, "boke.py")):
        pytest.skip("boke.py not in the reference install")
    t = lambda sd: {k: torch.from_numpy(np.asarray(a)) for k, a in sd.items()}
    torch.save({"model_state_dict": t(sd17)}, tmp_path / "policy.pt")
    torch.save({"model_state_dict": t(sd_value)}, tmp_path / "value.pt")
    cmd = [sys.executable, "-m", "bokego_b200.dropin", os.path.join(root, "boke.py"), "-g", "-t", "0.3",
           "-p", str(tmp_path / "policy.pt"), "-v", str(tmp_path / "value.pt")]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    p = subprocess.run(cmd, input="name\nplay b E5\ngenmove w\nshowboard\nquit\n", capture_output=True, text=True, timeout=300,
                       env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    replies = [r for r in p.stdout.split("\n\n") if r.strip()]
    assert replies[0].strip() == "= boke"
    assert replies[1].strip() == "="
    mv = replies[2].strip()[2:]
    after = go.Game()
    after.play_move(go.squash("E5"))
    assert after.is_legal(go.squash(mv)), replies
