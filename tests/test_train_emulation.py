"""CPU checks of the REINFORCE row's host logic and of the arithmetic the training kernels implement.

`emulate_*` below restate, in numpy, exactly what csrc/bk_train.cu computes -- the flat parameter layout (k = tap * Cin + ci),
the implicit-GEMM window gather, the data gradient as the same gather on dZ with mirrored taps and per-tap transposed
weights, the weight gradient, the per-position BatchNorm backward formula and the head / loss backward -- and compare the
result with the oracle (torch autograd over the reference's own ops).  The CUDA kernels themselves are compared with the
oracle on the GPU (tests/test_gpu_train.py); this file pins the conventions they share with the host code."""
import os

import numpy as np
import pytest
import torch

from bokego_b200 import reinforce as rf
from oracle import train as ot

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EPS = 1e-5


def taps(R):
    h = R // 2
    return [(t // R - h, t % R - h) for t in range(R * R)]


def gather(x, R, sign):
    """x [P,81,C] -> [P*81, R*R*C]: column (tap, ci) of row (p, sq) = x[p, sq shifted by sign * tap, ci], 0 off the board"""
    P, _, C = x.shape
    g = x.reshape(P, 9, 9, C)
    out = np.zeros((P, 9, 9, R * R, C), x.dtype)
    for t, (dx, dy) in enumerate(taps(R)):
        dx, dy = sign * dx, sign * dy
        xs, xe = max(0, -dx), min(9, 9 - dx)
        ys, ye = max(0, -dy), min(9, 9 - dy)
        out[:, xs:xe, ys:ye, t] = g[:, xs + dx:xe + dx, ys + dy:ye + dy]
    return out.reshape(P * 81, R * R * C)


def emulate(flat, planes, moves, coef):
    P = planes.shape[0]
    x0 = np.zeros((P, 81, 32), np.float64)
    x0[:, :, :27] = np.transpose(planes.astype(np.float64), (0, 2, 1))
    W = [flat[rf.TP_W0:rf.TP_W1].reshape(800, 128).astype(np.float64)] + \
        [flat[rf.TP_W1 + l * 147456: rf.TP_W1 + (l + 1) * 147456].reshape(1152, 128).astype(np.float64) for l in range(6)]
    vec = flat[rf.TP_VEC:rf.TP_HEADW].reshape(7, 3, 128).astype(np.float64)
    hw, hb = flat[rf.TP_HEADW:rf.TP_HEADW + 128].astype(np.float64), flat[rf.TP_HEADB:rf.TP_HEADB + 81].astype(np.float64)
    a, zs, acts, mus, rss = x0, [], [x0], [], []
    for l in range(7):
        R = 5 if l == 0 else 3
        z = (gather(a, R, +1) @ W[l] + vec[l, 0]).reshape(P, 81, 128)
        mu = z.mean(1, keepdims=True)
        rs = 1.0 / np.sqrt(z.var(1, keepdims=True) + EPS)
        a = np.maximum(0.0, (z - mu) * rs * vec[l, 1] + vec[l, 2])
        zs.append(z); acts.append(a); mus.append(mu); rss.append(rs)
    logits = a @ hw + hb
    e = np.exp(logits - logits.max(1, keepdims=True))
    pr = e / e.sum(1, keepdims=True)
    pm = pr[np.arange(P), moves]
    clamped = (pm < ot.PROB_EPS) | (pm > 1 - ot.PROB_EPS)        # Categorical clamps before the log: no gradient there
    nlp = -np.log(np.clip(pm, ot.PROB_EPS, 1 - ot.PROB_EPS))
    dl = np.where(clamped[:, None], 0.0, coef[:, None] * (pr - np.eye(81)[moves]))
    grad = np.zeros(rf.TP_COUNT, np.float64)
    grad[rf.TP_HEADB:rf.TP_HEADB + 81] = dl.sum(0)
    grad[rf.TP_HEADW:rf.TP_HEADW + 128] = np.einsum("ps,psc->c", dl, a)
    da = dl[:, :, None] * hw[None, None, :]
    gv = grad[rf.TP_VEC:rf.TP_HEADW].reshape(7, 3, 128)
    for l in range(6, -1, -1):
        R = 5 if l == 0 else 3
        g = np.where(acts[l + 1] > 0, da, 0.0)
        xh = (zs[l] - mus[l]) * rss[l]
        s1, s2 = g.sum(1, keepdims=True), (g * xh).sum(1, keepdims=True)
        dz = vec[l, 1] * rss[l] * (g - s1 / 81 - xh * s2 / 81)
        gv[l, 0], gv[l, 1], gv[l, 2] = dz.sum((0, 1)), s2.sum((0, 1)), s1.sum((0, 1))
        gw = gather(acts[l], R, +1).T @ dz.reshape(P * 81, 128)
        if l == 0:
            grad[rf.TP_W0:rf.TP_W1] = gw.ravel()
        else:
            grad[rf.TP_W1 + (l - 1) * 147456: rf.TP_W1 + l * 147456] = gw.ravel()
            wd = np.transpose(W[l].reshape(9, 128, 128), (0, 2, 1)).reshape(1152, 128)   # [(tap, co)][ci]
            da = (gather(dz, 3, -1) @ wd).reshape(P, 81, 128)
    return logits, nlp, grad


def test_flat_layout_round_trip():
    sd = dict(np.load(os.path.join(GOLD, "weights_policy_17.npz")))
    flat = rf.flat_from_tensors(lambda k: sd[k])
    assert flat.shape == (rf.TP_COUNT,)
    back = rf.tensors_from_flat(flat)
    assert list(back) == rf.param_keys() == ot.param_keys()
    for k, v in back.items():
        assert v.shape == sd[k].shape and np.array_equal(v, sd[k]), k
    w0 = flat[rf.TP_W0:rf.TP_W1].reshape(25, 32, 128)
    assert np.all(w0[:, 27:] == 0)
    assert w0[7, 3, 5] == sd["conv.0.weight"][5, 3, 1, 2]          # tap = kh * 5 + kw
    w3 = flat[rf.TP_W1 + 2 * 147456: rf.TP_W1 + 3 * 147456].reshape(9, 128, 128)
    assert w3[5, 17, 99] == sd["conv.9.weight"][99, 17, 1, 2]


def test_game_major_order():
    o = rf.game_major(3, 2)
    assert o.tolist() == [0, 2, 4, 1, 3, 5]


def test_kernel_arithmetic_matches_autograd():
    sd = dict(np.load(os.path.join(GOLD, "weights_policy_17.npz")))
    G = np.load(os.path.join(GOLD, "reinforce.npz"))
    planes = G["black3/calls"][[3, 40, 77, 120]]
    rng = np.random.default_rng(0)
    moves = rng.integers(0, 81, len(planes))
    coef = np.array([0.5, -0.25, 1.0, -1.0 / 3])
    flat = rf.flat_from_tensors(lambda k: sd[k])
    logits, nlp, grad = emulate(flat, planes, moves, coef)
    loss, grads, lo = ot.reinforce_grads(sd, planes.astype(np.float32), moves, coef)
    assert np.abs(logits - lo.numpy()).max() < 2e-3
    assert abs(float((nlp * coef).sum()) - loss) < 1e-3 * max(1.0, abs(loss))
    mine = rf.tensors_from_flat(grad.astype(np.float32))
    for k, g in grads.items():
        g = g.numpy()
        scale = np.abs(g).max()
        if k.endswith(".bias") and k.split(".")[1] in ("0", "3", "6", "9", "12", "15", "18"):
            assert np.abs(mine[k]).max() < 1e-4 and scale < 1e-4   # conv bias: cancelled by the mean subtraction
            continue
        assert np.abs(mine[k] - g).max() <= 2e-4 * scale + 1e-7, (k, np.abs(mine[k] - g).max(), scale)


def test_running_filter_and_adamw_against_reference_run():
    """oracle restatements of the running-average filter and of AdamW against the recorded reference iteration"""
    sd = dict(np.load(os.path.join(GOLD, "weights_policy_17.npz")))
    G = np.load(os.path.join(GOLD, "reinforce.npz"))
    stride = int(G["stride"])
    for tag in ("black3", "white2"):
        calls = G[f"{tag}/calls"].astype(np.float32)
        _, means, uv = ot.train_forward(sd, calls)
        rs = ot.running_stats(sd, means.numpy(), uv.numpy())
        for k, v in rs.items():
            ref = G[f"{tag}/post/{k}"]
            if k.endswith("tracked"):
                assert int(v) == int(ref)
            else:
                assert np.abs(v - ref).max() < 1e-4, k
        color, bs = int(G[f"{tag}/color"]), int(G[f"{tag}/bs"])
        rf_from = int(G[f"{tag}/replay_from"])
        pos = ot.replay_positions(G[f"{tag}/lengths"], color)
        assert len(pos) == len(calls) - rf_from
        mv = np.array([G[f"{tag}/moves"][g, j] for g, j in pos])
        coef = ot.reference_coef(G[f"{tag}/lengths"], G[f"{tag}/results"], color, bs)
        _, grads, _ = ot.reinforce_grads(sd, calls[rf_from:], mv, coef)
        for k, g in grads.items():
            a = g.numpy().ravel()
            ref, amax = G[f"{tag}/grad/{k}"], float(G[f"{tag}/grad/{k}/absmax"])
            s = a if a.size <= 4096 else a[::stride]
            if amax < 1e-5:
                assert np.abs(s).max() < 1e-5          # conv biases: round-off noise in the reference as well
                continue
            assert np.abs(s - ref).max() <= 1e-4 * amax, k
            assert abs(np.sqrt((a.astype(np.float64) ** 2).sum()) - float(G[f"{tag}/grad/{k}/l2"])) <= 1e-4 * float(G[f"{tag}/grad/{k}/l2"])
            p, _, _ = ot.adamw_step(sd[k], g.numpy(), np.zeros_like(sd[k]), np.zeros_like(sd[k]), 1)
            post = G[f"{tag}/post/{k}"]
            ps = p.numpy().ravel()
            ps = ps if ps.size <= 4096 else ps[::stride]
            # Adam's first step moves every element by lr * g / (|g| + eps): where |g| is at round-off level its sign,
            # and with it the element's update, is noise -- at most 2 * lr, and only on a handful of elements
            d = np.abs(ps - post)
            assert d.max() <= 2.1e-5 and (d > 1e-7).sum() <= max(3, 5e-3 * d.size), (k, d.max(), (d > 1e-7).mean())


def test_reinforce_argument_errors_need_no_device():
    """the mirror of `reinforce` keeps the reference's ValueError for a bad colour (bin/selfplay.py:84), raised before anything touches a device"""
    with pytest.raises(ValueError):
        rf.reinforce(None, None, None, "green")


def test_trainer_has_no_cpu_path():
    """no CPU fallback: the training state refuses a non-CUDA device"""
    from bokego_b200 import _lib
    sd = dict(np.load(os.path.join(GOLD, "weights_policy_17.npz")))
    with pytest.raises(_lib.BokegoB200Error):
        rf.PolicyTrainer(sd, torch.device("cpu"))


def test_training_symbols_exported():
    """every training entry point of include/bokego_b200.h is exported by the built library"""
    from bokego_b200 import _lib
    L = _lib.lib()
    for name in ("bk_train_param_count", "bk_train_workspace_bytes", "bk_train_launches", "bk_train_forward", "bk_train_backward",
                 "bk_train_running_stats", "bk_adamw_step"):
        assert hasattr(L, name), name
    assert L.bk_train_param_count() == rf.TP_COUNT
    # 576 positions: 23 on a B200, one more launch per 3x3 conv for its split tail
    assert L.bk_train_launches(0, 576, 5) in (17, 23) and L.bk_train_launches(0, 16, 5) == 23 and L.bk_train_launches(1, 576, 1) == 38
    assert L.bk_train_workspace_bytes(0) == 0 and L.bk_train_workspace_bytes(16) > 16 * 81 * 128 * 4 * 14


def _tf32_cut(x):
    """the tensor core's own cut of an fp32 word to TF32 (10 mantissa bits, toward zero) = csrc/bk_train_tc.cu tf32_cut"""
    return (np.asarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _tf32_rna(x):
    """cvt.rna.tf32.f32 (round to nearest, ties away) as the weight pack kernel applies it"""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64) + np.uint64(0x1000)
    return (u & np.uint64(0xFFFFE000)).astype(np.uint32).view(np.float32)


def _bf16_rn(x):
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))
    return (u & np.uint64(0xFFFF0000)).astype(np.uint32).view(np.float32)


def test_3xtf32_split_is_fp32_grade_and_bf16_low_parts_are_not():
    """the arithmetic of the tcgen05 training GEMMs restated in numpy (exact products and sums in float64, so only the SPLIT is
    measured): activations x = hi + lo with hi = the word cut to TF32 and lo = the exact remainder, cut again by the tensor core;
    weights w = rna(w) + rna(w - rna(w)); the product is hi*hi + hi*lo + lo*hi (csrc/bk_train_tc.cu: the N = 256 MMA
    A_hi x [B_hi | B_lo] and A_lo x B_hi).  What is dropped is lo*lo and the second cut: ~2^-21 per product, so a K = 1,152 dot product
    keeps fp32 accuracy.  With the low-order products on bf16 copies (the rejected BK_R3_LO_BF16 build, profiles/r02zz_lo_bf16.txt)
    every product carries 2^-20 .. 2^-18 instead: 1e-6 of the largest entry where the split leaves 1.4e-7, several times worse already
    in one dot product, which the data gradient's cancellation then amplifies past the fp32-grade bar of the recorded iteration."""
    rng = np.random.default_rng(5)
    x = np.maximum(rng.standard_normal((256, 1152)), 0).astype(np.float32)          # post-ReLU activations
    w = (0.05 * rng.standard_normal((1152, 128))).astype(np.float32)
    exact = x.astype(np.float64) @ w.astype(np.float64)
    xh = _tf32_cut(x)
    xl = _tf32_cut(x - xh)
    wh = _tf32_rna(w)
    wl = _tf32_rna(w - wh)
    assert np.array_equal(xh.astype(np.float64) + (x - xh).astype(np.float64), x.astype(np.float64))   # the remainder is exact
    f = lambda a: a.astype(np.float64)
    split = f(xh) @ f(wh) + f(xh) @ f(wl) + f(xl) @ f(wh)
    bf = f(xh) @ f(wh) + f(_bf16_rn(x)) @ f(_bf16_rn(w - wh)) + f(_bf16_rn(x - xh)) @ f(_bf16_rn(w))
    scale = np.abs(exact).max()
    e_split, e_bf = np.abs(split - exact).max() / scale, np.abs(bf - exact).max() / scale
    assert e_split < 2e-7, e_split                      # below fp32's own 6e-8 .. 1e-7 per rounding of the result
    assert e_bf > 4 * e_split and e_bf < 1e-4, (e_bf, e_split)


def test_conv3_work_item_schedule():
    """the persistent 3x3 training kernel's work items (csrc/bk_train_tc.cu: bk_train_conv3_schedule, host arithmetic): whole tiles of
    128 raster rows (100 per position), the tiles left over after the last full round of the SMs as four single-channel-group items
    each when they are few, one channel group per item for the small-batch forward"""
    import ctypes as C
    from bokego_b200 import _lib
    L = _lib.lib()

    def sched(P, n_sm=148, ksplit=1):
        out = (C.c_int * 4)()
        assert L.bk_train_conv3_schedule(P, n_sm, ksplit, out) == 0
        return list(out)

    assert sched(576) == [450, 444, 444 + 4 * 6, 148]           # 16 games x 36 moves: three rounds and six tiles split four ways
    assert sched(190) == [149, 148, 152, 148]
    assert sched(220) == [172, 148, 148 + 4 * 24, 148]          # 24 tail tiles: the most
    assert sched(221) == [173, 0, 173, 148]                     # 25: run whole
    assert sched(2048) == [1600, 0, 1600, 148]                  # 120 left over: too many to split
    assert sched(100) == [79, 0, 79, 79]                        # less than one round: one CTA per tile
    assert sched(148 * 128 // 100) == [148, 0, 148, 148]        # exactly one round
    assert sched(16, ksplit=4) == [13, 0, 52, 52]               # the 16-position forward of a self-play move
    assert sched(64, ksplit=4) == [50, 0, 200, 148]
    assert sched(576, n_sm=132) == [450, 0, 450, 132]           # 54 left over on 132 SMs: no split
    for P in range(1, 700, 7):                                   # every tile is covered exactly once
        n_tiles, full, items, ctas = sched(P)
        assert n_tiles == (100 * P + 127) // 128 and ctas == min(items, 148)
        assert items == (n_tiles if full == 0 else full + 4 * (n_tiles - full)) and full % 148 == 0 and full <= n_tiles
    out = (C.c_int * 4)()
    assert L.bk_train_conv3_schedule(0, 148, 1, out) == -1 and L.bk_train_conv3_schedule(16, 148, 2, out) == -1
