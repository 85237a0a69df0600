"""ctypes front-end of oracle/bk_oracle.c -- the CPU checker.  TEST INFRASTRUCTURE ONLY.

Nothing under bokego_b200/ may import this module; it is used by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
Each wrapper names the reference function it restates (see bk_oracle.c for file:line).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbk_oracle.so")

NONE, PASS = -2, -1


def build(force=False):
    src = os.path.join(_HERE, "bk_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libbk_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        i8p, i16p, u8p, f32p, f64p = (C.POINTER(C.c_int8), C.POINTER(C.c_int16), C.POINTER(C.c_uint8),
                                      C.POINTER(C.c_float), C.POINTER(C.c_double))
        L.bko_features_batch.argtypes = [i8p, i16p, i16p, i16p, u8p, u8p, u8p, u8p, C.c_int]
        L.bko_features_batch.restype = None
        L.bko_score_batch.argtypes = [i8p, C.c_double, f64p, C.c_int]
        L.bko_score_batch.restype = None
        L.bko_score.argtypes = [i8p, C.c_double]
        L.bko_score.restype = C.c_double
        L.bko_possible_eye.argtypes = [i8p, C.c_int]
        L.bko_possible_eye.restype = C.c_int
        L.bko_possible_ko.argtypes = [i8p, C.c_int]
        L.bko_possible_ko.restype = C.c_int
        L.bko_is_legal.argtypes = [i8p, C.c_int, C.c_int, C.c_int]
        L.bko_is_legal.restype = C.c_int
        L.bko_play.argtypes = [i8p, i16p, i16p, i16p, u8p, C.c_int, C.c_int]
        L.bko_play.restype = C.c_int
        L.bko_get_move_mcts.argtypes = [i8p, C.c_int, C.c_int, f32p, f32p, C.c_int, C.c_uint64, C.c_uint32,
                                        C.POINTER(C.c_int)]
        L.bko_get_move_mcts.restype = C.c_int
        L.bko_get_move_selfplay.argtypes = [i8p, C.c_int, C.c_int, f32p, f32p]
        L.bko_get_move_selfplay.restype = C.c_int
        L.bko_exp_draws.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, f32p]
        L.bko_exp_draws.restype = None
        L.bko_step_batch.argtypes = [i8p, i16p, i16p, i16p, u8p, u8p, f32p, f32p, C.c_int, C.c_uint64,
                                     C.c_uint32, C.c_int, C.c_int, i16p, C.c_int]
        L.bko_step_batch.restype = None
        _lib = L
    return _lib


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def features_batch(boards, ko, last, turn, libs_in=None):
    """nnet.features over a batch.  Returns (feats u8 [B,27,81], legal u8 [B,81], libs_out u8 [B,81])."""
    boards = _c(boards, np.int8).reshape(-1, 81)
    B = boards.shape[0]
    ko, last, turn = _c(ko, np.int16), _c(last, np.int16), _c(turn, np.int16)
    libs_in = None if libs_in is None else _c(libs_in, np.uint8).reshape(B, 81)
    feats = np.empty((B, 27, 81), np.uint8)
    legal = np.empty((B, 81), np.uint8)
    libs_out = np.empty((B, 81), np.uint8)
    lib().bko_features_batch(_p(boards, C.c_int8), _p(ko, C.c_int16), _p(last, C.c_int16), _p(turn, C.c_int16),
                             _p(libs_in, C.c_uint8), _p(feats, C.c_uint8), _p(legal, C.c_uint8),
                             _p(libs_out, C.c_uint8), B)
    return feats, legal, libs_out


def score_batch(boards, komi=5.5):
    """Game.score() over a batch -> float64 [B]."""
    boards = _c(boards, np.int8).reshape(-1, 81)
    out = np.empty(boards.shape[0], np.float64)
    lib().bko_score_batch(_p(boards, C.c_int8), komi, _p(out, C.c_double), boards.shape[0])
    return out


def possible_eye(board, s):
    board = _c(board, np.int8)
    return lib().bko_possible_eye(_p(board, C.c_int8), int(s))


def is_legal(board, ko, turn, s):
    board = _c(board, np.int8)
    return lib().bko_is_legal(_p(board, C.c_int8), int(ko), int(turn), int(s))


def play(board, ko, last, turn, move, libs=None, libs_fresh=False):
    """Game.play_move.  Returns (status, board, ko, last, turn, libs); status 0 ok / 1 ko / 2 not_empty / 3 suicide."""
    b = np.array(board, dtype=np.int8).copy()
    k, l, t = (np.array([v], np.int16) for v in (ko, last, turn))
    lb = None if libs is None else np.array(libs, dtype=np.uint8).copy()
    st = lib().bko_play(_p(b, C.c_int8), _p(k, C.c_int16), _p(l, C.c_int16), _p(t, C.c_int16),
                        _p(lb, C.c_uint8), int(bool(libs_fresh)), int(move))
    return st, b, int(k[0]), int(l[0]), int(t[0]), lb


def get_move_mcts(board, ko, turn, probs, q):
    """Go_MCTS.get_move with injected draws.  Returns (move, n_draws, probs_after)."""
    board = _c(board, np.int8)
    p = np.array(probs, dtype=np.float32).copy()
    q = _c(q, np.float32).reshape(-1, 81)
    nd = C.c_int(0)
    mv = lib().bko_get_move_mcts(_p(board, C.c_int8), int(ko), int(turn), _p(p, C.c_float), _p(q, C.c_float),
                                 q.shape[0], 0, 0, C.byref(nd))
    return mv, nd.value, p


def get_move_selfplay(board, ko, turn, probs, q):
    board = _c(board, np.int8)
    p, q = _c(probs, np.float32), _c(q, np.float32)
    return lib().bko_get_move_selfplay(_p(board, C.c_int8), int(ko), int(turn), _p(p, C.c_float), _p(q, C.c_float))


def exp_draws(seed, game, move, tr):
    q = np.empty(81, np.float32)
    lib().bko_exp_draws(int(seed), int(game), int(move), int(tr), _p(q, C.c_float))
    return q


def step_batch(boards, ko, last, turn, libs, done, probs, mode, max_turn, q_inj=None, seed=0, game0=0):
    """One playout step in place over numpy state arrays; returns moves int16 [B]."""
    B = boards.shape[0]
    for a, dt in ((boards, np.int8), (ko, np.int16), (last, np.int16), (turn, np.int16), (done, np.uint8)):
        assert a.dtype == dt and a.flags.c_contiguous
    probs = _c(probs, np.float32)
    qv = 0
    if q_inj is not None:
        q_inj = _c(q_inj, np.float32).reshape(B, -1, 81)
        qv = q_inj.shape[1]
    moves = np.empty(B, np.int16)
    lib().bko_step_batch(_p(boards, C.c_int8), _p(ko, C.c_int16), _p(last, C.c_int16), _p(turn, C.c_int16),
                         _p(libs, C.c_uint8), _p(done, C.c_uint8), _p(probs, C.c_float), _p(q_inj, C.c_float),
                         qv, int(seed), int(game0), int(mode), int(max_turn), _p(moves, C.c_int16), B)
    return moves
