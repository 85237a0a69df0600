"""Batched tree search over the B200 hot path (BASELINE configs[2]: "MCTS genmove with batched leaf
evaluation, virtual loss").

Same search rule as the reference's `MCTS` in its default `no_sim` mode (/root/reference/bokego/mcts.py:15-255):
  * PUCT selection (mcts.py:219-234): child score = -V[c]/N[c] + c_puct * P_parent[move] * sqrt(max(1, sum_c N[c])) / (1 + N[c]);
  * a leaf is expanded (all legal moves, never PASS, mcts.py:309-317) only once it has been visited more than `expand_thresh`
    times (mcts.py:172-183); the root is expanded immediately (mcts.py:153-157);
  * the leaf is scored by the value net, terminal or not (mcts.py:133-151), and the value alternates sign on the way up
    (mcts.py:208-217); `choose` takes the most visited child and re-roots (mcts.py:110-131).
What differs is the machinery: the tree is a set of flat arrays indexed by node id (no per-node Python objects, no dict
caches), positions of all nodes live in one device-resident pool, children are made by one `bk_make_moves` launch per
expansion, the descents / back-ups run in a native host-side core of the library, and leaves are evaluated in batches: up to
`leaf_batch` descents that need the device are parked under a virtual loss before one encoder + one policy/value launch scores
every new leaf (descents that end in an already evaluated leaf are backed up at once).  With `leaf_batch=1` the visit counts
are those of the sequential rule.

`no_sim=False` is the reference's `--simulate` mode (mcts.py:133-151, 195-217, boke.py:24-25): every rollout also plays its
leaf out to the end of the game with moves drawn from the policy net (Go_MCTS.find_random_child), the result -- Black's +-1
from Game.score(), negated when White is to move at the leaf -- is summed into Q on the way up with alternating sign, and
selection mixes the two estimates, ((1 - w) Q + w V) / N with w = value_net_weight (0.5; 0 without a value net).  Here the
`leaf_batch` parked leaves of a batch are played out TOGETHER on the device (bokego_b200.playout.run_playouts from the leaf
positions with their carried liberty caches: bk_forward + bk_playout_step_encode per move, bk_score at the end); the random
stream of playout number r is keyed (seed, r, turn, try), so a search is reproducible.  One difference from the reference is
deliberate: its get_move zeroes rejected moves in the CACHED distribution of the node (mcts.py:357), which leaks into later
PUCT priors of own-eye moves and into later playouts through the same position; here priors are never mutated.

Ties in the arg-max go to the lowest move index (the reference iterates a Python set, i.e. its tie-break is arbitrary).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, go
from .batched import MODE_MCTS, NONE, PASS, Positions, evaluate_positions, features_batch, make_moves, policy_value_batch  # noqa: F401

MAX_TURNS = 80   # mcts.py:13


class MCTS:
    """root: a bokego_b200.go.Game (or None for the empty board).  policy_net / value_net: PackedNet (or the nnet mirror's
    PolicyNet / ValueNet, whose packed blobs are used).  kwargs as in the reference: expand_thresh (100),
    exploration_weight (4.0), no_sim (True), value_net_weight (0.5 in --simulate mode with a value net); additionally
    leaf_batch (descents per evaluation batch, default 1), seed (random stream of the --simulate playouts) and device."""

    def __init__(self, root=None, policy_net=None, value_net=None, **kwargs):
        self.device = _lib.require_device(kwargs.get("device", "cuda"))
        self.expand_thresh = kwargs.get("expand_thresh", 100)
        self.exploration_weight = kwargs.get("exploration_weight", 4.0)
        self.leaf_batch = max(1, int(kwargs.get("leaf_batch", 1)))
        self.no_sim = kwargs.get("no_sim", True)
        self.seed = int(kwargs.get("seed", 0))
        self._bind_nets(policy_net, value_net)
        if self.no_sim:                                   # mcts.py:66-71
            self.value_net_weight = 1.0
        elif not self.has_value:
            self.value_net_weight = 0.0
        else:
            self.value_net_weight = kwargs.get("value_net_weight", 0.5)
        self.n_playouts = 0     # --simulate playouts made so far (= the game id of the next playout's random stream)
        self.n_evals = 0        # positions sent through the nets
        self.n_eval_batches = 0
        self.set_root(root)

    def _bind_nets(self, policy_net, value_net):
        if policy_net is None:
            raise TypeError("Missing required keywork argument: 'policy_net'")
        if value_net is None and self.no_sim:
            raise TypeError("Keyword argument 'value_net' is required for no simulation mode")
        pk = lambda n: None if n is None else (n._packed(self.device) if hasattr(n, "_packed") else n)
        self.policy, self.value = pk(policy_net), pk(value_net)
        self.has_value = self.value is not None

    def _net_outputs(self, sub):
        """(probs float32 [n,81], value float64 [n], legal uint8 [n,81], libs uint8 tensor [n,81]) of the positions `sub`:
        ONE launch (bk_forward_positions: planes on chip, policy + value)"""
        _, probs, val, out = evaluate_positions(sub, self.policy, self.value, want=("legal", "libs"))
        val = np.zeros(sub.B) if val is None else val.double().cpu().numpy()
        return probs.cpu().numpy(), val, out["legal"].cpu().numpy(), out["libs"]

    def _playout_results(self, leaves, first_id):
        """--simulate: play the positions of the nodes `leaves` out to the end (mcts.py:195-206) and return Black's +-1 per
        leaf (Go_MCTS.reward, mcts.py:330-338).  Playout j uses the random stream of game id first_id + j."""
        from .playout import MCTS_MAX_TURN, n_steps_for, run_playouts
        idx = torch.as_tensor(leaves, dtype=torch.long, device=self.device)
        p = self.pool
        pos = Positions(p.boards[idx].contiguous(), p.ko[idx].contiguous(), p.last[idx].contiguous(), p.turn[idx].contiguous(),
                        p.libs[idx].contiguous())
        pos.done = ((pos.turn > MAX_TURNS) | (pos.last == PASS)).to(torch.uint8)        # a terminal leaf is scored as it is
        first_turn = int(min(self.turn[i] for i in leaves))
        res = run_playouts(pos, self.policy, MODE_MCTS, MCTS_MAX_TURN, seed=self.seed, game0=first_id,
                           n_steps=n_steps_for(MODE_MCTS, MCTS_MAX_TURN, first_turn), first_turn=first_turn, graph=False)
        return res.reward.cpu().numpy().astype(np.float64)

    # ---- storage ---------------------------------------------------------------------------------------------------
    def _alloc(self, cap):
        dev = self.device
        self.cap = cap
        self.pool = Positions(torch.zeros(cap, 81, dtype=torch.int8, device=dev), torch.full((cap,), -1, dtype=torch.int16, device=dev),
                              torch.full((cap,), NONE, dtype=torch.int16, device=dev), torch.zeros(cap, dtype=torch.int16, device=dev),
                              torch.zeros(cap, 81, dtype=torch.uint8, device=dev))
        self.parent = np.full(cap, -1, np.int32)
        self.move = np.full(cap, NONE, np.int16)        # the move that led to the node
        self.N = np.zeros(cap, np.int64)
        self.V = np.zeros(cap, np.float64)
        self.Q = np.zeros(cap, np.float64)              # --simulate: playout reward sums (mcts.py:46)
        self.child0 = np.full(cap, -1, np.int32)        # first child id, children are contiguous
        self.nchild = np.full(cap, -1, np.int32)        # -1 = not expanded
        self.val = np.full(cap, np.nan, np.float64)     # value net output (mover's perspective), nan = not evaluated
        self.turn = np.zeros(cap, np.int16)
        self.last = np.full(cap, NONE, np.int16)
        self.prior = np.zeros((cap, 81), np.float32)    # policy probabilities of the node's position
        self.legal = np.zeros((cap, 81), np.uint8)
        self.n = 0

    def _grow(self, need):
        if self.n + need <= self.cap:
            return
        cap = max(self.cap * 2, self.n + need)
        old, n = self.pool, self.n
        keep = {k: getattr(self, k) for k in ("parent", "move", "N", "V", "Q", "child0", "nchild", "val", "turn", "last", "prior", "legal")}
        self._alloc(cap)
        self.n = n
        for k, v in keep.items():
            getattr(self, k)[: len(v)] = v
        for a, b in ((self.pool.boards, old.boards), (self.pool.ko, old.ko), (self.pool.last, old.last), (self.pool.turn, old.turn),
                     (self.pool.libs, old.libs)):
            a[: b.shape[0]] = b

    def set_root(self, root=None):
        """Start a new tree at `root` (a go.Game, or None = empty board): evaluate it and expand it (mcts.py:153-157)."""
        self._alloc(4096)
        if root is None:
            root = go.Game()
        lut = {go.BLACK: 1, go.WHITE: -1, go.EMPTY: 0}
        bd = torch.tensor([lut[c] for c in root.board], dtype=torch.int8)
        self.pool.boards[0] = bd.to(self.device)
        self.pool.ko[0] = -1 if root.ko is None else int(root.ko)
        last = PASS if root.last_move == go.PASS else (NONE if not isinstance(root.last_move, int) else int(root.last_move))
        self.pool.last[0] = last
        self.pool.turn[0] = int(root.turn)
        self.turn[0], self.last[0] = int(root.turn), last
        self._root_fresh = root._libs is None
        if root._libs is not None:
            self.pool.libs[0] = torch.frombuffer(bytearray(root._libs), dtype=torch.uint8).to(self.device)
        self.n = 1
        self.root = 0
        self._evaluate([0])
        self._expand(0)

    # ---- evaluation and expansion -------------------------------------------------------------------------------------
    def _evaluate(self, ids):
        """value, priors, legal moves and the refreshed liberty cache for the nodes `ids` (one encoder + one forward launch)"""
        ids = [i for i in dict.fromkeys(ids) if np.isnan(self.val[i])]
        if not ids:
            return
        idx = torch.as_tensor(ids, dtype=torch.long, device=self.device)
        fresh = ids == [0] and getattr(self, "_root_fresh", False) and self.parent[0] < 0
        sub = Positions(self.pool.boards[idx].contiguous(), self.pool.ko[idx].contiguous(), self.pool.last[idx].contiguous(),
                        self.pool.turn[idx].contiguous(), None if fresh else self.pool.libs[idx].contiguous())
        probs, val, legal, libs = self._net_outputs(sub)
        self.pool.libs[idx] = libs
        self.prior[ids] = probs
        self.val[ids] = np.asarray(val, np.float64)
        self.legal[ids] = legal
        self.n_evals += len(ids)
        self.n_eval_batches += 1

    def _terminal(self, i):
        return self.turn[i] > MAX_TURNS or self.last[i] == PASS      # mcts.py:362-364

    def _expand(self, i):
        self._expand_many([i])

    def _expand_many(self, nodes):
        """children (all legal moves, never PASS, mcts.py:309-317) for every node of `nodes`: one bk_make_moves launch for all"""
        par, mvs, spans = [], [], []
        for i in nodes:
            if self.nchild[i] >= 0:
                continue
            if self._terminal(i):
                self.nchild[i] = 0
                continue
            mv = np.flatnonzero(self.legal[i]).astype(np.int16)
            self.nchild[i] = len(mv)
            if len(mv):
                spans.append((i, len(mv)))
                par.append(np.full(len(mv), i, np.int32))
                mvs.append(mv)
        if not spans:
            return
        par, mvs = np.concatenate(par), np.concatenate(mvs)
        c = len(mvs)
        self._grow(c)
        lo = self.n
        child, _ = make_moves(self.pool, torch.from_numpy(par).to(self.device), torch.from_numpy(mvs).to(self.device))
        sl = slice(lo, lo + c)
        self.pool.boards[sl], self.pool.ko[sl], self.pool.last[sl] = child.boards, child.ko, child.last
        self.pool.turn[sl], self.pool.libs[sl] = child.turn, child.libs
        self.parent[sl], self.move[sl] = par, mvs
        self.turn[sl], self.last[sl] = self.turn[par] + 1, mvs
        at = lo
        for i, k in spans:
            self.child0[i] = at
            at += k
        self.n += c

    # ---- search -----------------------------------------------------------------------------------------------------------
    def rollout(self, n=1):
        """n rollouts.  The descents, the virtual losses and the back-ups run in the library's host-side tree core
        (bk_tree_run / bk_tree_finish, csrc/bk_tree.cu: mcts.py:172-234 on the flat arrays); it hands control back whenever
        up to `leaf_batch` descents wait for the device -- an unevaluated leaf, or a leaf visited more than expand_thresh times
        that has to get its children -- and those are evaluated / expanded here in one batch."""
        L = _lib.lib()
        K, D = self.leaf_batch, 128
        pend_nodes, pend_len = np.empty((K, D), np.int32), np.empty(K, np.int32)
        pend_expand, n_pend = np.empty(K, np.int32), C.c_int(0)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        sim, have_value = not self.no_sim, int(self.has_value)
        done = 0
        while done < n:
            # the arrays may have been re-allocated by an expansion: take the pointers afresh every round
            done += L.bk_tree_run(p(self.N), p(self.V), p(self.child0), p(self.nchild), p(self.move), p(self.prior), p(self.val),
                                  int(self.root), n - done, K, int(self.expand_thresh), C.c_double(self.exploration_weight),
                                  p(pend_nodes), p(pend_len), p(pend_expand), D, C.byref(n_pend),
                                  p(self.Q) if sim else None, C.c_double(self.value_net_weight), have_value)
            k = n_pend.value
            if k == 0:
                continue
            leaves = [int(pend_nodes[j, pend_len[j] - 1]) for j in range(k)]
            self._evaluate(leaves)           # priors and legal moves are needed with or without a value net
            self._expand_many(list(dict.fromkeys(int(x) for x in pend_expand[:k] if x >= 0)))
            reward = None
            if sim:
                # mcts.py:199-204: Black's result, seen by the player to move at the leaf
                black = self._playout_results(leaves, self.n_playouts)
                self.n_playouts += k
                reward = np.ascontiguousarray(np.where(self.turn[leaves] % 2 == 0, black, -black), np.float64)
            L.bk_tree_finish(p(self.N), p(self.V), p(self.val), p(pend_nodes), p(pend_len), k, D, K,
                             p(self.Q) if sim else None, None if reward is None else p(reward), have_value)
            done += k

    def root_visits(self):
        """visit counts of the root's children as an 81-vector indexed by move"""
        out = np.zeros(81, np.int64)
        lo, c = self.child0[self.root], max(0, self.nchild[self.root])
        out[self.move[lo: lo + c]] = self.N[lo: lo + c]
        return out

    def best_move(self):
        lo, c = self.child0[self.root], max(0, self.nchild[self.root])
        if c == 0:
            return PASS
        return int(self.move[lo + int(np.argmax(self.N[lo: lo + c]))])

    def choose(self):
        """most visited child of the root becomes the new root (mcts.py:110-131); returns its move"""
        lo, c = self.child0[self.root], max(0, self.nchild[self.root])
        if c == 0:
            return PASS
        best = lo + int(np.argmax(self.N[lo: lo + c]))
        self.root = best
        self._evaluate([best])
        self._expand(best)
        return int(self.move[best])

    def advance(self, mv):
        """Re-root at the child of the root reached by move `mv`, keeping the statistics of its subtree (what the reference's
        input_move -> set_root does, gtp.py:332-337: its N / V dicts keep every node).  False when the root has no such child
        (a pass, or a root that was never expanded): the caller then starts a new tree."""
        lo, c = self.child0[self.root], max(0, self.nchild[self.root])
        hit = np.flatnonzero(self.move[lo: lo + c] == mv)
        if len(hit) == 0:
            return False
        self.root = lo + int(hit[0])
        self._evaluate([self.root])
        self._expand(self.root)
        return True

    def winrate(self, node=None):
        """(((1 - w) Q + w V) / N + 1) / 2 from the perspective of the player to move at the node (mcts.py:159-170)"""
        i = self.root if node is None else node
        w = self.value_net_weight
        return (((1 - w) * self.Q[i] + w * self.V[i]) / self.N[i] + 1) / 2 if self.N[i] > 0 else 0
