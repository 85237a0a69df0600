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
    # the record kernel (bk_pack_records) against the plain tensor formula
    head = torch.stack([a.n_moves.to(torch.int16), a.reward.to(torch.int16), torch.round(a.score * 2).to(torch.int16)], dim=1)
    assert torch.equal(a.records(), torch.cat([head, a.moves], dim=1))
    g = po.PlayoutGraph(B, dev, p17, mode, seed=11, game0=5, policy_odd=odd).replay()
    torch.cuda.synchronize()
    assert torch.equal(g.records(), a.records())
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


@pytest.mark.parametrize("mode", [0, 1])
def test_fused_step_encode_equals_step_then_encode(env, mode):
    """bk_playout_step_encode == bk_playout_step followed by bk_encode (carried cache, in place): boards, ko / last / turn,
    liberty cache, done flags, moves and every byte of the conv operand, move after move to the end of the games; and the
    planes of the fused path decode to the oracle's features of the same positions"""
    bk, po, dev, p17, p19 = env
    B, seed = 203, 21
    max_turn = 80 if mode == 0 else 70
    a, b = bk.Positions.empty(B, dev), bk.Positions.empty(B, dev)
    fa = bk.features_batch(a, fresh_libs=True, want=("conv", "libs"), out={"libs": a.libs})
    fb = bk.features_batch(b, fresh_libs=True, want=("conv", "libs"), out={"libs": b.libs})
    for k in range(po.n_steps_for(mode, max_turn)):
        net = p17 if (mode == 0 or k % 2 == 0) else p19
        _, probs, _ = bk.policy_value_batch(fa["conv"], B, net, None, want_logits=False)
        ma = bk.playout_step(a, probs, mode, max_turn, seed=seed, game0=9, encode_into=fa["conv"])
        live = b.done == 0
        mb = bk.playout_step(b, probs, mode, max_turn, seed=seed, game0=9)
        fresh = bk.features_batch(b, fresh_libs=False, want=("conv", "libs"), out={"libs": b.libs})
        # boards that finished with this move keep their old planes in the fused path (nobody evaluates them again)
        still = (b.done == 0)
        conv_b = torch.where(_board_mask(still, fb["conv"].numel(), dev), fresh["conv"], fb["conv"])
        fb["conv"] = conv_b
        assert torch.equal(ma, mb), k
        for x, y in ((a.boards, b.boards), (a.ko, b.ko), (a.last, b.last), (a.turn, b.turn), (a.done, b.done)):
            assert torch.equal(x, y), k
        assert torch.equal(a.libs[still], b.libs[still]), k
        assert torch.equal(fa["conv"], fb["conv"]), k
        if k in (0, 7, 40) and bool(still.any()):
            idx = torch.nonzero(still).flatten()[:32]
            sub = bk.Positions(a.boards[idx].contiguous(), a.ko[idx].contiguous(), a.last[idx].contiguous(), a.turn[idx].contiguous())
            want, _, _ = ocpu.features_batch(*(t.cpu().numpy() for t in (sub.boards, sub.ko, sub.last, sub.turn)), None)
            got = _decode_conv(fa["conv"], B)[idx.cpu().numpy()]
            # planes 6..12 (liberties) depend on the carried cache; every other plane is a function of the position alone
            keep = [c for c in range(27) if not 6 <= c <= 12]
            assert np.array_equal(got[:, keep], want[:, keep]), k
        if not bool(live.any()):
            break
    assert bool((a.done != 0).all())


def _board_mask(board_flags, nbytes, dev):
    """byte mask over the conv operand: True for the bytes of boards whose flag is set ([group][4 chunks][605 rows][16 B])"""
    B = board_flags.numel()
    G = (B + 4) // 5
    flags = torch.zeros(G * 5, dtype=torch.bool, device=dev)
    flags[:B] = board_flags
    rows = flags.reshape(G, 5, 1).expand(G, 5, 121).reshape(G, 1, 605, 1).expand(G, 4, 605, 16)
    return rows.reshape(-1)[:nbytes]


def _decode_conv(conv, B):
    """conv operand -> uint8 planes [B][27][81]"""
    G = (B + 4) // 5
    h = conv.view(torch.float16).reshape(G, 4, 605, 8).cpu().float().numpy()
    out = np.zeros((G * 5, 27, 81), np.uint8)
    for bi in range(5):
        for x in range(9):
            for y in range(9):
                r = 121 * bi + 22 + 11 * x + y
                v = h[:, :, r, :].reshape(G, 32)[:, :27]
                out[bi::5][:, :, 9 * x + y] = v.astype(np.uint8)
    return out[:B]


def test_simulate_65536_boards_subset_vs_oracle(env):
    """BASELINE configs[3] at its full size: 65,536 concurrent --simulate playouts.  The CUDA-graph run equals the launch-by-launch
    run record for record; and 96 boards spread over the batch are replayed by the oracle, move for move, from the
    probabilities the device computed for them (bit-exact playouts given identical probabilities and draws), down to the
    final boards, scores and rewards."""
    bk, po, dev, p17, _ = env
    B, seed, game0 = 65536, 31, 1000
    a = po.run_playouts(bk.Positions.empty(B, dev, track_libs=False), p17, bk.MODE_MCTS, seed=seed, game0=game0, graph=True)
    rec_a = a.records()
    pick = np.unique(np.concatenate([np.arange(0, B, 701), [B - 1, B - 2, 4, 5, 739, 740]]))[:96]
    pt = torch.from_numpy(pick).to(dev)
    pos = bk.Positions.empty(B, dev)
    bufs = bk.features_batch(pos, fresh_libs=True, want=("conv", "libs"), out={"libs": pos.libs})
    probs = torch.empty(B, 81, dtype=torch.float32, device=dev)
    n = len(pick)
    bd = np.zeros((n, 81), np.int8); ko = np.full(n, -1, np.int16); last = np.full(n, -2, np.int16)
    turn = np.zeros(n, np.int16); done = np.zeros(n, np.uint8)
    _, _, libs = ocpu.features_batch(bd, ko, last, turn, None)
    moves = []
    for k in range(po.n_steps_for(bk.MODE_MCTS, 80)):
        bk.policy_value_batch(bufs["conv"], B, p17, None, want_logits=False, probs_out=probs)
        sub = probs[pt].cpu().numpy()
        mv = bk.playout_step(pos, probs, bk.MODE_MCTS, 80, seed=seed, game0=game0, encode_into=bufs["conv"])
        moves.append(mv.clone())
        # the oracle plays the picked boards with the device's probabilities; its random stream is keyed by the GLOBAL game id
        mo = np.empty(n, np.int16)
        for j, gid in enumerate(pick):
            sl = slice(j, j + 1)
            mo[j] = ocpu.step_batch(bd[sl], ko[sl], last[sl], turn[sl], libs[sl], done[sl], sub[sl], 0, 80, seed=seed,
                                    game0=game0 + int(gid))[0]
        assert np.array_equal(mv[pt].cpu().numpy(), mo), k
        _, _, libs = ocpu.features_batch(bd, ko, last, turn, libs)
    assert bool(pos.done.all()) and done.all()
    assert np.array_equal(pos.boards[pt].cpu().numpy(), bd) and np.array_equal(pos.turn[pt].cpu().numpy(), turn)
    score, reward = bk.score_batch(pos.boards)
    assert np.array_equal(score[pt].cpu().numpy().astype(np.float64), ocpu.score_batch(bd))
    b_rec = po.PlayoutResult(torch.stack(moves, 1), pos.turn, score, reward).records()
    assert torch.equal(rec_a, b_rec)
    assert set(np.unique(reward.cpu().numpy())) <= {-1, 1}


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("B", [1, 7, 203, 512, 745, 1485])
def test_persistent_playout_kernel_equals_launch_per_move(env, mode, B):
    """bk_playout_run (whole games inside one launch of the conv kernel: policy forward, then three warps per board sample,
    play and re-encode into the shared-memory operand of the next move) against the launch-per-move loop: every move, the
    final boards / ko / last / turn / liberty caches / done flags, scores and rewards must be identical.  The item size adapts
    to B (1 board per item up to 148 boards, 2 at 203, 4 at 512, 5 from 741 on); 745 ends in a CTA pair whose second CTA has no
    boards, 1485 needs more than one round of the 74 CTA pairs."""
    bk, po, dev, p17, p19 = env
    odd = p19 if mode == 1 else None
    a_pos, b_pos = bk.Positions.empty(B, dev, track_libs=False), bk.Positions.empty(B, dev, track_libs=False)
    a = po.run_playouts(a_pos, p17, mode, seed=13, game0=77, policy_odd=odd, persistent=True)
    b = po.run_playouts(b_pos, p17, mode, seed=13, game0=77, policy_odd=odd, persistent=False, graph=False)
    torch.cuda.synchronize()
    assert torch.equal(a.moves, b.moves)
    for x, y in ((a_pos.boards, b_pos.boards), (a_pos.ko, b_pos.ko), (a_pos.last, b_pos.last), (a_pos.turn, b_pos.turn),
                 (a_pos.done, b_pos.done), (a_pos.libs, b_pos.libs), (a.score, b.score), (a.reward, b.reward)):
        assert torch.equal(x, y)
    assert bool(a_pos.done.all())


def test_persistent_playout_from_midgame_positions_and_repeats(env):
    """playouts that start from arbitrary positions with a carried liberty cache (the leaves of a --simulate search), odd
    first turn, repeated launches with cold and warm L2: identical to the launch-per-move loop every time"""
    bk, po, dev, p17, p19 = env
    B = 333
    base = bk.Positions.empty(B, dev)
    uni = torch.full((B, 81), 1.0 / 81, dtype=torch.float32, device=dev)
    for _ in range(23):                                   # 23 random legal moves: White to move, captures and kos on the boards
        bk.features_batch(base, fresh_libs=False, want=("libs",), out={"libs": base.libs})
        bk.playout_step(base, uni, bk.MODE_MCTS, 1000, seed=4, game0=0)
    keep = [t.clone() for t in (base.boards, base.ko, base.last, base.turn, base.libs)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    want = None
    for rep in range(6):
        pos = bk.Positions(*(t.clone() for t in keep))
        pos.done = ((pos.turn > 80) | (pos.last == -1)).to(torch.uint8)
        if rep % 2:
            flush.zero_()
        res = po.run_playouts(pos, p17, bk.MODE_MCTS, seed=3, game0=5, first_turn=23, n_steps=po.n_steps_for(bk.MODE_MCTS, 80, 23),
                              persistent=rep > 0, graph=False)
        torch.cuda.synchronize()
        got = (res.moves.clone(), pos.boards.clone(), pos.libs.clone(), res.score.clone())
        if want is None:
            want = got
        for x, y in zip(got, want):
            assert torch.equal(x, y), rep


@pytest.mark.parametrize("n_steps", [1, 10])
def test_persistent_playout_stopped_early_leaves_the_same_state(env, n_steps):
    """a run that stops before the games are over: positions, liberty caches and done flags equal the launch-per-move loop's, so
    either engine can continue from the other's state"""
    bk, po, dev, p17, p19 = env
    B = 300
    a_pos, b_pos = bk.Positions.empty(B, dev, track_libs=False), bk.Positions.empty(B, dev, track_libs=False)
    a = po.run_playouts(a_pos, p17, bk.MODE_SELFPLAY, seed=2, game0=0, policy_odd=p19, n_steps=n_steps, persistent=True)
    b = po.run_playouts(b_pos, p17, bk.MODE_SELFPLAY, seed=2, game0=0, policy_odd=p19, n_steps=n_steps, persistent=False, graph=False)
    assert torch.equal(a.moves, b.moves) and not bool(a_pos.done.any())
    for x, y in ((a_pos.boards, b_pos.boards), (a_pos.ko, b_pos.ko), (a_pos.last, b_pos.last), (a_pos.turn, b_pos.turn),
                 (a_pos.done, b_pos.done), (a_pos.libs, b_pos.libs)):
        assert torch.equal(x, y)
    # ... and continue crosswise to the end
    a2 = po.run_playouts(a_pos, p17, bk.MODE_SELFPLAY, seed=2, game0=0, policy_odd=p19, first_turn=n_steps,
                         n_steps=po.n_steps_for(bk.MODE_SELFPLAY, 70, n_steps), persistent=False, graph=False)
    b2 = po.run_playouts(b_pos, p17, bk.MODE_SELFPLAY, seed=2, game0=0, policy_odd=p19, first_turn=n_steps,
                         n_steps=po.n_steps_for(bk.MODE_SELFPLAY, 70, n_steps), persistent=True)
    assert torch.equal(a2.moves, b2.moves) and torch.equal(a_pos.boards, b_pos.boards) and torch.equal(a2.score, b2.score)
