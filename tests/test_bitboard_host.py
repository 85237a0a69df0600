"""The kernels' bit-board rules (bk_bitboard.cuh), compiled for the host, against the golden
vectors and the C oracle.  CPU only: catches rule bugs before any GPU time is spent."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import cpu as ocpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hb(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hb") / "libhb.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++",
                           os.path.join(HERE, "host_bitboard.cpp"), "-o", so])
    L = C.CDLL(so)
    i8p, u8p = C.POINTER(C.c_int8), C.POINTER(C.c_uint8)
    L.hb_features.argtypes = [i8p, C.c_int, C.c_int, C.c_int, u8p, u8p, u8p, u8p]
    L.hb_play.argtypes = [i8p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    L.hb_is_legal.argtypes = [i8p, C.c_int, C.c_int, C.c_int]
    L.hb_eye.argtypes = [i8p, C.c_int]
    L.hb_score_diff.argtypes = [i8p]
    L.hb_exp_draws.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def test_features_vs_golden(hb, positions):
    p = positions
    n = len(p["board"])
    for i in range(0, n, 2):
        bd = np.ascontiguousarray(p["board"][i])
        libs_in = None if p["fresh"][i] else np.ascontiguousarray(p["libs_in"][i])
        f = np.zeros((27, 81), np.uint8); lg = np.zeros(81, np.uint8); lo = np.zeros(81, np.uint8)
        hb.hb_features(_p(bd, C.c_int8), int(p["ko"][i]), int(p["last"][i]), int(p["turn"][i]),
                       None if libs_in is None else _p(libs_in, C.c_uint8), _p(f, C.c_uint8), _p(lg, C.c_uint8),
                       _p(lo, C.c_uint8))
        assert np.array_equal(f, p["feats"][i]), i
        assert np.array_equal(lg, p["legal"][i]) and np.array_equal(lo, p["libs_out"][i]), i


def test_rules_vs_golden(hb, positions, rules):
    p, r = positions, rules
    for j, i in enumerate(r["src"]):
        bd0 = np.ascontiguousarray(p["board"][i])
        ko, last, turn = int(p["ko"][i]), int(p["last"][i]), int(p["turn"][i])
        for s in range(81):
            bd = bd0.copy()
            k, l, t = C.c_int(ko), C.c_int(last), C.c_int(turn)
            st = hb.hb_play(_p(bd, C.c_int8), C.byref(k), C.byref(l), C.byref(t), s)
            assert st == r["status"][j][s]
            assert np.array_equal(bd, r["nboard"][j][s])
            if st == 0:
                assert k.value == r["nko"][j][s] and l.value == s and t.value == turn + 1
            assert hb.hb_is_legal(_p(bd0, C.c_int8), ko, turn, s) == r["islegal"][j][s]
            assert hb.hb_eye(_p(bd0, C.c_int8), s) == r["eye"][j][s]
        assert hb.hb_score_diff(_p(bd0, C.c_int8)) - 5.5 == r["score"][j]


def test_random_boards_vs_oracle(hb):
    """boards that never occur in play (random fill, random ko) must agree with the oracle too"""
    rng = np.random.default_rng(3)
    for it in range(300):
        dens = rng.uniform(0.1, 0.95)
        bd = rng.choice(np.array([0, 1, -1], np.int8), size=81, p=[1 - dens, dens / 2, dens / 2]).astype(np.int8)
        empt = np.where(bd == 0)[0]
        ko = int(rng.choice(empt)) if len(empt) and rng.random() < 0.5 else -1
        last = int(rng.integers(-2, 81)); turn = int(rng.integers(0, 90))
        libs_in = rng.integers(0, 9, 81).astype(np.uint8) if it % 2 else None
        fo, lgo, loo = ocpu.features_batch(bd[None], [ko], [last], [turn], None if libs_in is None else libs_in[None])
        f = np.zeros((27, 81), np.uint8); lg = np.zeros(81, np.uint8); lo = np.zeros(81, np.uint8)
        hb.hb_features(_p(bd, C.c_int8), ko, last, turn, None if libs_in is None else _p(libs_in, C.c_uint8),
                       _p(f, C.c_uint8), _p(lg, C.c_uint8), _p(lo, C.c_uint8))
        assert np.array_equal(f, fo[0]) and np.array_equal(lg, lgo[0]) and np.array_equal(lo, loo[0]), it
        assert hb.hb_score_diff(_p(bd, C.c_int8)) - 5.5 == ocpu.score_batch(bd[None])[0]
        for s in range(81):
            assert hb.hb_is_legal(_p(bd, C.c_int8), ko, turn, s) == ocpu.is_legal(bd, ko, turn, s)
            assert hb.hb_eye(_p(bd, C.c_int8), s) == ocpu.possible_eye(bd, s)


def test_exp_stream_bit_identical(hb):
    for seed, g, m, t in ((0, 0, 0, 0), (12345678901234, 77, 13, 5), (2**63 + 5, 4095, 80, 81)):
        q = np.zeros(81, np.float32)
        hb.hb_exp_draws(seed, g, m, t, _p(q, C.c_float))
        assert np.array_equal(q, ocpu.exp_draws(seed, g, m, t))
