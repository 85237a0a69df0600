"""Pins the CPU oracle (oracle/bk_oracle.c, oracle/nets.py) to vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import cpu as ocpu
from oracle import nets as onets


def test_features_all_golden_positions(positions):
    p = positions
    fresh = p["fresh"].astype(bool)
    for sel, libs in ((fresh, None), (~fresh, p["libs_in"][~fresh])):
        f, legal, lo = ocpu.features_batch(p["board"][sel], p["ko"][sel], p["last"][sel], p["turn"][sel], libs)
        assert np.array_equal(f, p["feats"][sel])
        assert np.array_equal(legal, p["legal"][sel])
        assert np.array_equal(lo, p["libs_out"][sel])
    assert fresh.sum() > 2000 and (~fresh).sum() > 2000


def test_sgf_and_quirk_positions_present(positions):
    tags = set(positions["tag"].tolist())
    assert set(range(1, 11)) <= tags and {200, 201, 202, 203, 204, 205} <= tags
    # F5: the 3-stone group touching the played point twice is counted 6 -> plane 25 (value 6)
    i = int(np.where(positions["tag"] == 200)[0][0])
    assert positions["feats"][i][25][20] == 6


def test_rules_play_legal_eye_score(positions, rules):
    p, r = positions, rules
    for j, i in enumerate(r["src"]):
        bd, ko, last, turn = p["board"][i], int(p["ko"][i]), int(p["last"][i]), int(p["turn"][i])
        for s in range(81):
            st, nb, nko, nlast, nturn, _ = ocpu.play(bd, ko, last, turn, s)
            assert st == r["status"][j][s], (i, s)
            if st == 0:
                assert np.array_equal(nb, r["nboard"][j][s]) and nko == r["nko"][j][s]
                assert nlast == s and nturn == turn + 1
            else:
                assert np.array_equal(nb, bd)
            assert ocpu.is_legal(bd, ko, turn, s) == r["islegal"][j][s]
            assert ocpu.possible_eye(bd, s) == r["eye"][j][s]
        st, nb, nko, nlast, nturn, _ = ocpu.play(bd, ko, last, turn, -1)
        assert (nko, nlast, nturn) == tuple(r["pass_state"][j]) and np.array_equal(nb, bd)
    assert np.array_equal(ocpu.score_batch(p["board"][r["src"]]), r["score"])


def test_known_scores(positions, rules):
    # 21.5 and 3.5 (border overwrite quirk)
    src = rules["src"].tolist()
    for tag, want in ((201, 21.5), (202, 3.5)):
        i = int(np.where(positions["tag"] == tag)[0][0])
        assert rules["score"][src.index(i)] == want


def test_nets_restatement(positions, nets_golden, sd17, sd19, sd_value):
    x = onets.planes_to_float(positions["feats"][nets_golden["src"]])
    torch.set_num_threads(1)
    assert float((onets.policy_logits(sd17, x) - torch.from_numpy(nets_golden["logits17"])).abs().max()) < 1e-4
    assert float((onets.policy_logits(sd19, x) - torch.from_numpy(nets_golden["logits19"])).abs().max()) < 1e-4
    assert float((onets.value(sd_value, x) - torch.from_numpy(nets_golden["value"])).abs().max()) < 1e-5


def test_mcts_flavour_moves(playouts):
    t = playouts
    n = len(t["m_move"])
    for i in range(n):
        mv, nd, _ = ocpu.get_move_mcts(t["m_board"][i], t["m_ko"][i], t["m_turn"][i], t["m_probs"][i], t["m_q"][i])
        assert mv == t["m_move"][i] and nd == t["m_nq"][i], i
    assert (t["m_move"] == -1).sum() >= 1 and t["m_nq"].max() > 10


def test_mcts_flavour_whole_playouts(playouts):
    """step the oracle from the empty board with the recorded probs/draws: same moves, boards, reward"""
    t = playouts
    for g in np.unique(t["m_game"]):
        idx = np.where(t["m_game"] == g)[0]
        bd = np.zeros((1, 81), np.int8); ko = np.array([-1], np.int16); last = np.array([-2], np.int16)
        turn = np.zeros(1, np.int16); done = np.zeros(1, np.uint8); libs = np.zeros((1, 81), np.uint8)
        for i in idx:
            assert np.array_equal(bd[0], t["m_board"][i]) and turn[0] == t["m_turn"][i] and not done[0]
            f, _, lo = ocpu.features_batch(bd, ko, last, turn, None if t["m_fresh"][i] else libs)
            libs[:] = lo
            mv = ocpu.step_batch(bd, ko, last, turn, libs, done, t["m_probs"][i][None], 0, 80, q_inj=t["m_q"][i][None])
            assert mv[0] == t["m_move"][i]
        assert done[0] and np.array_equal(bd[0], t["m_final_board"][g]) and turn[0] == t["m_final_turn"][g]
        sc = ocpu.score_batch(bd)[0]
        assert sc == t["m_score"][g] and (1 if sc > 0 else -1) == t["m_reward"][g]


def test_selfplay_flavour(playouts):
    t = playouts
    for g in np.unique(t["s_game"]):
        idx = np.where(t["s_game"] == g)[0]
        bd = np.zeros((1, 81), np.int8); ko = np.array([-1], np.int16); last = np.array([-2], np.int16)
        turn = np.zeros(1, np.int16); done = np.zeros(1, np.uint8); libs = np.zeros((1, 81), np.uint8)
        for i in idx:
            assert np.array_equal(bd[0], t["s_board"][i]) and turn[0] == t["s_turn"][i] and not done[0]
            _, _, lo = ocpu.features_batch(bd, ko, last, turn, None if t["s_fresh"][i] else libs)
            if not t["s_fresh"][i]:
                assert np.array_equal(libs[0], t["s_libs_in"][i])
            libs[:] = lo
            mv = ocpu.step_batch(bd, ko, last, turn, libs, done, t["s_probs"][i][None], 1, 70, q_inj=t["s_q"][i][None])
            assert mv[0] == t["s_move"][i]
        assert done[0] and np.array_equal(bd[0], t["s_final_board"][g]) and turn[0] == t["s_final_turn"][g]
        assert (1 if ocpu.score_batch(bd)[0] > 0 else -1) == t["s_result"][g]
    assert (t["s_sampled_legal"] == 0).sum() > 10


def test_exp_stream_is_exponential():
    q = np.stack([ocpu.exp_draws(42, g, 3, 0) for g in range(2000)]).ravel().astype(np.float64)
    assert q.min() > 0 and abs(q.mean() - 1.0) < 0.01 and abs(q.var() - 1.0) < 0.03
    assert abs(np.mean(q > 1.0) - np.exp(-1.0)) < 0.005
    # different counters give different streams; same counters repeat
    assert np.array_equal(ocpu.exp_draws(1, 2, 3, 4), ocpu.exp_draws(1, 2, 3, 4))
    assert not np.array_equal(ocpu.exp_draws(1, 2, 3, 4), ocpu.exp_draws(1, 2, 3, 5))


def test_standin_value_head_fixture_matches_generator():
    """bench.py's own arm loads the stand-in value head from a fixture (it must not import oracle/): same numbers"""
    import os
    from oracle import nets as onets
    fx = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "weights_value_head_standin.npz")))
    gen = onets.standin_value_head(1234)
    assert set(fx) == set(gen)
    for k, v in gen.items():
        assert np.array_equal(fx[k], v.numpy()), k


def test_f16_operand_emulation_is_a_bounded_perturbation_of_the_reference_nets(positions, nets_golden, sd17):
    """oracle.nets.policy_logits_f16_operands (the yardstick of tests/test_gpu_precision.py: conv weights and inter-layer
    activations rounded to fp16, fp32 sums) stays within the rounding floor of 16-bit operands of the reference's fp32 logits on the
    golden positions, keeps their arg-max, and is not simply the fp32 result"""
    import torch
    from oracle import nets as onets
    src = nets_golden["src"][:128]
    x = onets.planes_to_float(positions["feats"][src])
    want = torch.from_numpy(nets_golden["logits17"][:128])
    exact = onets.policy_logits(sd17, x)
    emu = onets.policy_logits_f16_operands(sd17, x)
    assert float((exact - want).abs().max()) < 1e-4
    err = float((emu - want).abs().max())
    assert 1e-3 < err < 5e-2, err
    assert bool((emu.argmax(1) == want.argmax(1)).all())
    assert float((torch.softmax(emu, 1) - torch.softmax(want, 1)).abs().max()) < 2e-3
