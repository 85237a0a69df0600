"""REINFORCE row on the device (csrc/bk_train.cu through the C ABI) against the oracle (oracle/train.py: torch CPU autograd over
the reference's own ops) and against the iteration recorded from the unmodified reference (tests/golden/reinforce.npz).

Precisions: 5 = tcgen05 3xTF32 (default), 4 = tcgen05 TF32, 1 / 0 = the same two on warp-level mma.sync, 2 = FFMA validation path.
Tolerances (float32 storage everywhere), gradients measured per tensor against the tensor's largest entry:
  3xTF32 (5, 1) and FFMA (2): logits 2e-3 abs; gradients 1e-3.  Typical agreement is 3e-6; the limit is set by ReLU: a
      pre-activation within round-off of zero takes the other branch in one of the two implementations, and the gradient of
      that one unit (one term among thousands in every weight-gradient entry) appears or disappears.
  TF32 (4, 0; opt-in fast modes: operands cut to 10 mantissa bits -- what cuDNN does by default for the reference on a
      GPU): logits 5e-2 abs with the same arg-max wherever the margin allows; gradients within 0.2 per tensor and cosine
      similarity of the whole gradient >= 0.99 (rounding errors compound through 14 GEMMs and the per-position
      normalisation's 1/sigma).  Mode 4 hands raw fp32 words to the tensor core, which TRUNCATES them (a biased error,
      where mode 0 rounds to nearest): logits 0.25, gradients 0.3.
  The recorded reference iteration (108 positions, the coefficient on 36 of them) is held to 5e-3 at the worst entry and to
  1e-4 at the 90th percentile of every tensor: ReLU flips touch few entries, everything else agrees to fp32 accuracy.
  conv biases: their true gradient under per-position BatchNorm is zero (the mean subtraction cancels them); the reference
  and the kernels both produce round-off noise there (|g| < 1e-4 against 0.1 .. 10 elsewhere).
"""
import os

import numpy as np
import pytest
import torch

from oracle import nets as onets
from oracle import train as ot

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONV_BIAS = {f"conv.{i}.bias" for i in (0, 3, 6, 9, 12, 15, 18)}
TOL = {0: (5e-2, 0.2), 1: (2e-3, 1e-3), 2: (2e-3, 1e-3), 4: (0.25, 0.3), 5: (2e-3, 1e-3)}      # prec -> (logits abs, gradient relative to the largest entry)


@pytest.fixture(scope="module")
def env(sd17):
    from bokego_b200 import reinforce as rf
    dev = torch.device("cuda", 0)
    G = dict(np.load(os.path.join(GOLD, "reinforce.npz")))
    return rf, dev, G


def _dev(a, dev, dt):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)


REPORT = {}


def _check_grads(mine, want, rel, what=""):
    worst = 0.0
    for k, g in want.items():
        g = g.numpy() if isinstance(g, torch.Tensor) else g
        m = mine[k].numpy()
        scale = float(np.abs(g).max())
        if k in CONV_BIAS and scale < 1e-4:
            assert np.abs(m).max() < 1e-4, (what, k)
            continue
        err = float(np.abs(m - g).max())
        worst = max(worst, err / scale)
        REPORT[f"{what} {k}"] = err / scale
    REPORT[f"{what} WORST"] = worst
    print(f"[grad error] {what}: worst {worst:.3e} of the tensor's largest entry (limit {rel:.0e})")
    bad = {k: v for k, v in REPORT.items() if k.startswith(what + " ") and v > rel}
    assert not bad, bad


@pytest.mark.parametrize("prec", [2, 1, 0, 5, 4])
@pytest.mark.parametrize("P", [1, 37, 130])
def test_train_mode_forward(env, sd17, prec, P):
    """PolicyNet.forward in train() mode, one position per call (nnet.py:265-275): logits, SOFT probabilities and the
    per-position statistics BatchNorm feeds into its running averages"""
    rf, dev, G = env
    planes = G["black3/calls"][:P]
    tr = rf.PolicyTrainer(sd17, dev, prec=prec)
    logits, probs, stats = tr.forward(_dev(planes, dev, torch.uint8), rf.BN_POSITION, want_probs=True, want_stats=True)
    want, means, uvars = ot.train_forward(sd17, planes.astype(np.float32))
    lg = logits.cpu()
    assert float((lg - want).abs().max()) < TOL[prec][0]
    if prec not in (0, 4):
        assert bool((lg.argmax(1) == want.argmax(1)).all())
    else:
        top2 = want.topk(2, dim=1).values
        safe = (top2[:, 0] - top2[:, 1]) > 2 * TOL[prec][0]
        assert bool((lg.argmax(1) == want.argmax(1))[safe].all())
    assert float((probs.cpu() - torch.softmax(want, 1)).abs().max()) < ({0: 2e-2, 4: 6e-2}.get(prec, 1e-3))
    assert float((probs.cpu().sum(1) - 1).abs().max()) < 1e-5
    st = stats.cpu()
    tol = {0: 3e-2, 4: 8e-2}.get(prec, 1e-3)
    assert float(((st[:, :, 0] - means).abs() / (1 + means.abs())).max()) < tol
    assert float(((st[:, :, 1] - uvars).abs() / (1 + uvars.abs())).max()) < tol


@pytest.mark.parametrize("prec", [2, 0, 5, 4])
def test_eval_mode_forward(env, sd17, prec):
    """bn_mode 1 = running statistics: the same function as the fused inference kernel / PolicyNet.eval()"""
    rf, dev, G = env
    planes = G["black3/calls"][:50]
    tr = rf.PolicyTrainer(sd17, dev, prec=prec)
    logits, _, _ = tr.forward(_dev(planes, dev, torch.uint8), rf.BN_EVAL)
    want = onets.policy_logits(sd17, planes.astype(np.float32).reshape(-1, 27, 9, 9))
    assert float((logits.cpu() - want).abs().max()) < TOL[prec][0]


@pytest.mark.parametrize("prec", [2, 1, 0, 5, 4])
@pytest.mark.parametrize("bn", ["position", "eval"])
def test_gradients_vs_autograd(env, sd17, prec, bn):
    rf, dev, G = env
    planes = G["white2/calls"][5:5 + 45]
    rng = np.random.default_rng(3)
    moves = rng.integers(0, 81, len(planes)).astype(np.int16)
    coef = rng.uniform(-1, 1, len(planes)).astype(np.float32)
    coef[7] = 0.0
    tr = rf.PolicyTrainer(sd17, dev, prec=prec)
    mode = rf.BN_POSITION if bn == "position" else rf.BN_EVAL
    tr.forward(_dev(planes, dev, torch.uint8), mode)
    nlp = tr.backward(_dev(moves, dev, torch.int16), _dev(coef, dev, torch.float32))
    loss, grads, logits = ot.reinforce_grads(sd17, planes.astype(np.float32), moves, coef, bn)
    want_nlp = -ot.log_prob(logits, torch.from_numpy(moves))
    assert float((nlp.cpu() - want_nlp).abs().max()) < ({0: 5e-2, 4: 0.25}.get(prec, 2e-3))
    _check_grads(tr.grads_dict(), grads, TOL[prec][1], f"prec {prec} bn {bn}")
    flat_want = torch.from_numpy(rf.flat_from_tensors(lambda k: grads[k])).double()
    cos = float(torch.dot(tr.grads.cpu().double(), flat_want) / (tr.grads.cpu().double().norm() * flat_want.norm()))
    print(f"[grad cosine] prec {prec} bn {bn}: {cos:.6f}")
    assert cos >= (0.99 if prec in (0, 4) else 0.999999)
    # deterministic: a second pass gives the same bits; accumulate adds
    g1 = tr.grads.clone()
    tr.forward(_dev(planes, dev, torch.uint8), mode)
    tr.backward(_dev(moves, dev, torch.int16), _dev(coef, dev, torch.float32))
    assert torch.equal(g1, tr.grads)
    tr.backward(_dev(moves, dev, torch.int16), _dev(coef, dev, torch.float32), accumulate=True)
    assert float((tr.grads - 2 * g1).abs().max()) <= 1e-5 * float(g1.abs().max())


@pytest.mark.parametrize("P", [190, 200, 220, 576])
def test_conv3_split_tail_equals_whole_tiles(env, sd17, P):
    """the persistent 3x3 kernel's work items at the sizes where the schedule changes (148 SMs; a tile = 128 rows of the 100-row-per-position
    raster): 190 positions = 149 tiles = one full round of the grid and a split tail of 1 tile (4 single-channel-group items whose
    partial results a small kernel adds), 200 -> 9 tail tiles, 220 -> 24 (the most a tail may have), 576 -> 3 rounds + 6.  The same
    kernel with the tail switched off (BK_TC_NO_TAIL, read per launch) runs those tiles whole: only the order of four partial sums
    differs, so logits and every gradient tensor agree to round-off (measured 2e-7 .. 5e-7: profiles/r02w_tail_check.txt).  Against
    the FFMA path both differ by the same 1e-3 .. 9e-3 -- ReLU branches of pre-activations within round-off of zero -- hence this
    pairing and not that one."""
    rf, dev, G = env
    calls = np.concatenate([G["black3/calls"], G["white2/calls"]])
    planes = _dev(calls[np.arange(P) % len(calls)], dev, torch.uint8)
    rng = np.random.default_rng(P)
    moves = _dev(rng.integers(0, 81, P), dev, torch.int16)
    coef = _dev(rng.uniform(-1, 1, P), dev, torch.float32)
    out = []
    try:
        for off in (False, True):
            if off:
                os.environ["BK_TC_NO_TAIL"] = "1"
            tr = rf.PolicyTrainer(sd17, dev, prec=5)
            logits = tr.forward(planes)[0].clone()
            tr.backward(moves, coef)
            out.append((logits, tr.grads_dict()))
    finally:
        os.environ.pop("BK_TC_NO_TAIL", None)
    assert float((out[0][0] - out[1][0]).abs().max()) <= 2e-6 * float(out[1][0].abs().max())
    differs = False
    for k, want in out[1][1].items():
        w, m = want.numpy(), out[0][1][k].numpy()
        amax = float(np.abs(w).max())
        differs = differs or not np.array_equal(w, m)
        if amax >= 1e-4:                                           # (conv biases in front of a BatchNorm: zero up to round-off)
            assert np.abs(m - w).max() <= 2e-6 * amax, (P, k, float(np.abs(m - w).max() / amax))
    assert differs, "the split tail was not exercised (the two runs are bit-identical)"


def test_clamped_log_prob_has_no_gradient(env, sd17):
    """Categorical clamps probabilities to [eps, 1 - eps] before the log: a move below eps costs -log(eps) and has zero gradient"""
    rf, dev, G = env
    planes = G["black3/calls"][:8]
    logits, _, _ = ot.train_forward(sd17, planes.astype(np.float32))
    p = torch.softmax(logits, 1)
    moves = p.argmin(1).numpy().astype(np.int16)
    assert float(p.min(1).values.max()) < ot.PROB_EPS          # every chosen move is below the clamp
    tr = rf.PolicyTrainer(sd17, dev, prec=2)
    tr.forward(_dev(planes, dev, torch.uint8))
    nlp = tr.backward(_dev(moves, dev, torch.int16), torch.ones(8, device=dev))
    assert float((nlp.cpu() + np.log(ot.PROB_EPS)).abs().max()) < 1e-5
    assert float(tr.grads.abs().max()) == 0.0


def test_chunked_step_equals_single_pass(env, sd17):
    rf, dev, G = env
    planes = _dev(G["black3/calls"][:100], dev, torch.uint8)
    rng = np.random.default_rng(5)
    moves = _dev(rng.integers(0, 81, 100), dev, torch.int16)
    coef = _dev(rng.uniform(-1, 1, 100), dev, torch.float32)
    a, b = rf.PolicyTrainer(sd17, dev), rf.PolicyTrainer(sd17, dev)
    la = rf.reinforce_step(a, planes, moves, coef, chunk=100)
    lb = rf.reinforce_step(b, planes, moves, coef, chunk=32)
    assert abs(float(la) - float(lb)) < 1e-3 * max(1.0, abs(float(la)))
    assert float((a.grads - b.grads).abs().max()) <= 1e-4 * float(a.grads.abs().max())
    assert float((a.params - b.params).abs().max()) <= 2.1e-5


@pytest.mark.parametrize("tag", ["black3", "white2"])
@pytest.mark.parametrize("prec", [5, 1, 2, 0, 4])
def test_reference_iteration(env, sd17, tag, prec):
    """one iteration of the unmodified reference's `reinforce` (bin/selfplay.py:59-122): its stream of train-mode calls, the
    games it played and their results go in; p.grad, the parameters after AdamW and the running statistics must come out"""
    rf, dev, G = env
    stride = int(G["stride"])
    color, bs = int(G[f"{tag}/color"]), int(G[f"{tag}/bs"])
    calls, rfrom = G[f"{tag}/calls"], int(G[f"{tag}/replay_from"])
    lengths, results, moves = G[f"{tag}/lengths"], G[f"{tag}/results"], G[f"{tag}/moves"]
    pos = ot.replay_positions(lengths, color)
    mv = np.array([moves[g, j] for g, j in pos], np.int16)
    coef = ot.reference_coef(lengths, results, color, bs)
    tr = rf.PolicyTrainer(sd17, dev, prec=prec)
    # running statistics: every train-mode call in order (self-play calls, then the replay calls)
    _, _, stats = tr.forward(_dev(calls, dev, torch.uint8), want_stats=True)
    tr.update_running(stats)
    # the step itself on the replayed positions
    planes = _dev(calls[rfrom:], dev, torch.uint8)
    loss = rf.reinforce_step(tr, planes, _dev(mv, dev, torch.int16), _dev(coef, dev, torch.float32))
    want_loss, _, _ = ot.reinforce_grads(sd17, calls[rfrom:].astype(np.float32), mv, coef)
    assert abs(float(loss) - want_loss) < ({0: 5e-2, 4: 0.25}.get(prec, 2e-3)) * max(1.0, abs(want_loss))
    grads, post = tr.grads_dict(), tr.state_dict()
    rel = max(TOL[prec][1], 5e-3)
    for k in rf.param_keys():
        ref, amax = G[f"{tag}/grad/{k}"], float(G[f"{tag}/grad/{k}/absmax"])
        a = grads[k].numpy().ravel()
        s = a if a.size <= 4096 else a[::stride]
        if amax < 1e-5:
            assert np.abs(s).max() < 1e-4, k
        else:
            e = np.abs(s - ref)
            assert e.max() <= rel * amax, (k, e.max(), amax)
            if prec not in (0, 4):       # fp32-grade modes: almost every entry agrees to 1e-4; ReLU flips touch a few
                assert np.quantile(e, 0.9) <= 1e-4 * amax, (k, np.quantile(e, 0.9), amax)
            l2 = float(np.sqrt((a.astype(np.float64) ** 2).sum()))
            assert abs(l2 - float(G[f"{tag}/grad/{k}/l2"])) <= rel * float(G[f"{tag}/grad/{k}/l2"]), k
        # AdamW's first step moves an element by lr * g / (|g| + eps): +-lr wherever |g| >> eps, so parameters agree to a
        # few 1e-8 except where a round-off-level gradient changes sign (at most 2 * lr)
        p = post[k].numpy().ravel()
        p = p if p.size <= 4096 else p[::stride]
        d = np.abs(p - G[f"{tag}/post/{k}"])
        assert d.max() <= 2.1e-5, k
        if k not in CONV_BIAS and prec not in (0, 4):
            assert (d > 1e-7).mean() < 0.02, (k, (d > 1e-7).mean())
    for i in (1, 4, 7, 10, 13, 16, 19):
        for w in ("running_mean", "running_var"):
            ref = G[f"{tag}/post/conv.{i}.{w}"]
            err = np.abs(post[f"conv.{i}.{w}"].numpy() - ref) / (1 + np.abs(ref))
            assert err.max() < ({0: 3e-2, 4: 8e-2}.get(prec, 1e-4)), (i, w, err.max())
        assert int(post[f"conv.{i}.num_batches_tracked"]) == int(G[f"{tag}/post/conv.{i}.num_batches_tracked"])


def test_adamw_three_steps(env):
    rf, dev, G = env
    from bokego_b200 import _lib
    import ctypes as C
    L = _lib.lib()
    rng = np.random.default_rng(9)
    n = 5000
    p0 = rng.normal(size=n).astype(np.float32)
    p, m, v = _dev(p0, dev, torch.float32), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    wp, wm, wv = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step in (1, 2, 3):
        g = (rng.normal(size=n) * 10.0 ** rng.integers(-6, 2, n)).astype(np.float32)
        rc = L.bk_adamw_step(_lib.ptr(p), _lib.ptr(_dev(g, dev, torch.float32)), _lib.ptr(m), _lib.ptr(v), C.c_size_t(n),
                             C.c_double(1e-3), C.c_double(0.9), C.c_double(0.999), C.c_double(1e-8), C.c_double(0.01), step,
                             _lib.stream_ptr(dev))
        assert rc == 0
        wp, wm, wv = (t.numpy() for t in ot.adamw_step(wp, g, wm, wv, step, lr=1e-3))
    assert np.abs(p.cpu().numpy() - wp).max() < 1e-6
    assert np.allclose(m.cpu().numpy(), wm, rtol=1e-5, atol=1e-12) and np.allclose(v.cpu().numpy(), wv, rtol=1e-5, atol=1e-20)


@pytest.mark.parametrize("color", ["black", "white"])
def test_reinforce_loop(env, sd17, sd19, color):
    """the mirror of `reinforce` end to end: self-play on the device with `pi` in train mode, the step, the hand-back into the
    torch module and optimizer; and what it computed is re-derived by the oracle from the games it played"""
    rf, dev, G = env
    from bokego_b200 import nnet
    pi, opp = nnet.PolicyNet(), nnet.PolicyNet()
    pi.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd17.items()})
    opp.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd19.items()})
    pi.to(dev).train()
    opp.to(dev).eval()
    opt = torch.optim.AdamW(pi.parameters(), lr=1e-5)
    before = {k: v.detach().cpu().clone() for k, v in pi.state_dict().items()}
    stats = []
    tr = rf.reinforce(pi, opp, opt, color, n_itrs=2, bs=4, device=dev, stats=stats, id=0, seed=5, prec=rf.PREC_FFMA,
                      accumulate="batch")
    assert len(stats) == 2 and all(0 <= w <= 4 for w in stats)
    after = pi.state_dict()
    for k in rf.param_keys():
        d = float((after[k].cpu() - before[k]).abs().max())
        assert 0 < d <= 2.3e-5, (k, d)                 # two AdamW steps of lr 1e-5
    assert int(after["conv.1.num_batches_tracked"]) == int(before["conv.1.num_batches_tracked"]) + 2 * 2 * 36 * 4
    assert float((after["conv.1.running_mean"].cpu() - before["conv.1.running_mean"]).abs().max()) > 0
    st = opt.state[dict(pi.named_parameters())["conv.3.weight"]]
    assert int(st["step"]) == 2 and st["exp_avg"].shape == (128, 128, 3, 3)
    # the trained module still evaluates through the fused inference kernel
    pi.eval()
    x = torch.zeros(2, 27, 9, 9, device=dev)
    assert pi(x).shape == (2, 81)
    # the recorded positions of the last iteration, differentiated by the oracle with the parameters before that step, are
    # not available any more (parameters moved on) -- but the positions themselves must be real self-play positions:
    planes = tr._rec_planes.cpu().numpy()
    assert planes.shape == (36, 4, 27, 81)
    assert (planes[0, :, 0].sum(1) == 0).all()                 # plane 0 = stones of the player to move: none at pi's first move
    assert (planes[0, :, 1].sum(1) == (1 if color == "white" else 0)).all()   # plane 1 = opponent stones
    assert (planes[-1, :, 2].sum(1) < 50).all()                # plane 2 = empty squares, at the training colour's last move


def test_properties_at_scale(env, sd17):
    """size-independent properties on a batch the oracle would take minutes for (1,024 positions, two forward/backward chunks):
    the gradient is linear in the coefficients -- scaling them by 2 scales every entry by exactly 2 (a power of two: bit for bit) --
    positions with a zero coefficient contribute nothing, and the order of the positions only changes the order of the sums"""
    rf, dev, G = env
    P = 1024
    calls = np.concatenate([G["black3/calls"], G["white2/calls"]])
    planes = _dev(calls[np.arange(P) % len(calls)], dev, torch.uint8)
    rng = np.random.default_rng(11)
    moves = _dev(rng.integers(0, 81, P), dev, torch.int16)
    coef = _dev(rng.uniform(-1, 1, P), dev, torch.float32)
    coef[rng.integers(0, P, 300)] = 0.0
    tr = rf.PolicyTrainer(sd17, dev)
    rf.compute_grads(tr, planes, moves, coef, chunk=512)
    g1 = tr.grads.clone()
    assert bool(torch.isfinite(g1).all()) and float(g1.abs().max()) > 0
    rf.compute_grads(tr, planes, moves, 2 * coef, chunk=512)
    assert torch.equal(tr.grads, 2 * g1)
    keep = torch.nonzero(coef != 0).reshape(-1)
    rf.compute_grads(tr, planes[keep].contiguous(), moves[keep].contiguous(), coef[keep].contiguous(), chunk=512)
    scale = float(g1.abs().max())
    assert float((tr.grads - g1).abs().max()) <= 2e-5 * scale
    perm = torch.from_numpy(rng.permutation(P)).to(dev)
    rf.compute_grads(tr, planes[perm].contiguous(), moves[perm].contiguous(), coef[perm].contiguous(), chunk=512)
    assert float((tr.grads - g1).abs().max()) <= 2e-5 * scale
