"""Host-side tree core of the library (csrc/bk_tree.cu) in --simulate mode against a line-by-line Python restatement of the
reference's selection and back-up (mcts.py:208-234): paths, N, V and Q must agree exactly.  No GPU needed."""
import ctypes as C
from math import sqrt

import numpy as np

from bokego_b200 import _lib


def test_sim_mode_selection_and_backup():
    L = _lib.lib()
    rng = np.random.default_rng(0)
    n = 1 + 5 + 20                           # root, 5 children, 4 grandchildren each (terminal: nchild = 0)
    N, V, Q = np.zeros(n, np.int64), np.zeros(n), np.zeros(n)
    child0, nchild, move = np.full(n, -1, np.int32), np.full(n, -1, np.int32), np.full(n, -2, np.int16)
    prior, val = rng.random((n, 81)).astype(np.float32), rng.uniform(-1, 1, n)
    child0[0], nchild[0], move[1:6] = 1, 5, np.arange(5)
    for c in range(5):
        child0[1 + c], nchild[1 + c], move[6 + 4 * c: 10 + 4 * c] = 6 + 4 * c, 4, np.arange(4)
    nchild[6:] = 0
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    K, D, w = 1, 16, 0.5
    pn, pl, pe, npend = np.empty((K, D), np.int32), np.empty(K, np.int32), np.empty(K, np.int32), C.c_int(0)
    rew = rng.choice([-1.0, 1.0], 200)
    N2, V2, Q2 = N.copy(), V.copy(), Q.copy()

    def select(i):                           # mcts.py:219-234
        lo, c = child0[i], nchild[i]
        total = max(1, sum(N2[lo: lo + c]))
        best, best_s = None, -1e300
        for ch in range(lo, lo + c):
            avg = 0 if N2[ch] == 0 else ((1 - w) * Q2[ch] + w * V2[ch]) / N2[ch]
            s = -avg + 4.0 * float(prior[i, move[ch]]) * sqrt(total) / (1 + N2[ch])
            if s > best_s:
                best, best_s = ch, s
        return best

    for r in range(200):
        done = L.bk_tree_run(p(N), p(V), p(child0), p(nchild), p(move), p(prior), p(val), 0, 1, K, 1000, C.c_double(4.0), p(pn), p(pl),
                             p(pe), D, C.byref(npend), p(Q), C.c_double(w), 1)
        assert done == 0 and npend.value == 1            # --simulate: every descent waits for its playout
        path = [0]
        while nchild[path[-1]] > 0:
            path.append(select(path[-1]))
        assert list(pn[0, : pl[0]]) == path, r
        L.bk_tree_finish(p(N), p(V), p(val), p(pn), p(pl), 1, D, K, p(Q), p(np.array([rew[r]])), 1)
        v, q = val[path[-1]], rew[r]
        for nd in reversed(path):                        # mcts.py:208-217
            N2[nd] += 1; Q2[nd] += q; q = -q; V2[nd] += v; v = -v
    assert np.array_equal(N, N2) and np.array_equal(V, V2) and np.array_equal(Q, Q2)


def test_virtual_loss_is_taken_back():
    L = _lib.lib()
    n = 1 + 8
    N, V, Q = np.zeros(n, np.int64), np.zeros(n), np.zeros(n)
    child0, nchild, move = np.full(n, -1, np.int32), np.full(n, 0, np.int32), np.arange(-1, 8).astype(np.int16)
    child0[0], nchild[0] = 1, 8
    prior, val = np.full((n, 81), 1 / 81, np.float32), np.linspace(-0.5, 0.5, n)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    K, D = 4, 8
    pn, pl, pe, npend = np.empty((K, D), np.int32), np.empty(K, np.int32), np.empty(K, np.int32), C.c_int(0)
    L.bk_tree_run(p(N), p(V), p(child0), p(nchild), p(move), p(prior), p(val), 0, 4, K, 1000, C.c_double(4.0), p(pn), p(pl), p(pe), D,
                  C.byref(npend), p(Q), C.c_double(0.5), 1)
    assert npend.value == 4 and N[0] == 4 and Q[0] == 4.0 and len({int(pn[j, 1]) for j in range(4)}) == 4   # four different leaves
    rew = np.array([1.0, -1.0, 1.0, 1.0])
    L.bk_tree_finish(p(N), p(V), p(val), p(pn), p(pl), 4, D, K, p(Q), p(rew), 1)
    assert N[0] == 4 and N[1:].sum() == 4 and Q[0] == -rew.sum() and abs(V[0] + sum(val[int(pn[j, 1])] for j in range(4))) < 1e-12
