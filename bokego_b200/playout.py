"""Whole playouts and self-play games on the device, sharded over ranks (SURVEY 8b `playout_batch`, 8e).

Whole playouts are ONE kernel launch (bk_playout_run, `persistent=True`; the default while every board gets an SM in one round): the conv kernel keeps each item of up to
five boards on its SM for the whole game, positions resident on chip -- three warps per board encode the starting position,
and after every policy forward they sample / play / capture / re-encode straight into the shared-memory operand of the next
move -- followed by one bk_score.  The launch-per-move form (`persistent=False`) does the same with two launches per move, boards resident in HBM:
    bk_forward (policy only)  ->  bk_playout_step_encode (sample, play, capture, re-encode the new position in place)
and is what a net in training uses (reinforce.PolicyTrainer records planes and statistics move by move).  Both give the same
games, bit for bit.  It is exactly the loop of the reference's `MCTS._simulate` (/root/reference/bokego/mcts.py:195-206: `find_random_child`
until terminal, then `reward`) and of `bin/selfplay.py:18-33` (`playout`: `legal_sample` for pi_1 / pi_2 alternately), run for
all boards at once.  The steps of a whole game are captured once in a CUDA graph and replayed (the per-step work is small
at self-play batch sizes, so launch latency matters).

Games are independent: rank r of R owns the contiguous block of global game ids shard_range(n, r, R); the random stream is
keyed by the GLOBAL id, so results do not depend on R.  The only communication is one all_gather of fixed-size records.
"""
import torch

from . import _lib
from .batched import (MODE_MCTS, MODE_SELFPLAY, Positions, features_batch, pack_records, playout_run, playout_step,
                      policy_value_batch, score_batch)

MCTS_MAX_TURN = 80        # mcts.py:13  (terminal iff turn > 80 or the move was PASS, mcts.py:362-364)
SELFPLAY_MAX_TURN = 70    # bin/selfplay.py:16 (checked every two moves => 72 moves, selfplay.py:21-33)


def shard_range(n, rank, world):
    """contiguous block [lo, hi) of n units owned by `rank` of `world` (sizes differ by at most one)"""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class PlayoutResult:
    """moves int16 [B, T] (BK_MOVE_* codes once a board is finished), n_moves int16 [B] (turn reached),
    score float32 [B] (Game.score()), reward int8 [B] (+1 black wins, -1 otherwise)"""

    def __init__(self, moves, n_moves, score, reward, records=None):
        self.moves, self.n_moves, self.score, self.reward = moves, n_moves, score, reward
        self._records = records

    def records(self):
        """fixed-size per-game records int16 [B, T + 3]: n_moves, reward, 2*score, moves..."""
        if self._records is not None:
            return self._records
        if self.moves.is_cuda:      # one launch (bk_pack_records) from the [T, B] move log
            return pack_records(self.moves.t().contiguous(), self.n_moves.to(torch.int16).contiguous(), self.score, self.reward)
        head = torch.stack([self.n_moves.to(torch.int16), self.reward.to(torch.int16),
                            torch.round(self.score * 2).to(torch.int16)], dim=1)
        return torch.cat([head, self.moves], dim=1).contiguous()


def record_to_sgf(record, out_path, **kwargs):
    """One game record (a row of PlayoutResult.records(): n_moves, reward, 2*score, moves...) -> an SGF file written by
    go.write_sgf (the reference's record format, /root/reference/bokego/go.py:538-582; read back by go.get_moves).
    Moves after the end of the game (BK_MOVE_* codes below PASS) are dropped; the result is B+/W+ from the reward."""
    from . import go
    rec = [int(v) for v in record]
    moves = [m for m in rec[3:] if m >= go.PASS]
    kwargs.setdefault("result", ("B+" if rec[1] > 0 else "W+") + str(abs(rec[2]) / 2))
    return go.write_sgf(moves, out_path, **kwargs)


def persistent_pays(B, device):
    """The persistent playout kernel keeps an item of boards on its SM for the whole game, so its step phase (sample, play,
    re-encode: ~10 us per move) is serial with that item's policy evaluation, while the launch-per-move loop amortises one
    stepping launch over all rounds of the grid.  Measured on the B200 (profiles/r02k_playout_per_move.jsonl,
    r02l_playout_per_move.jsonl): the kernel wins while the boards fit two rounds of 4-board items (64 boards: 47 against 54 us
    per move; 512: 85 against 104; 1,024: 170 against 182), ties at 1,480 (212 / 208) and loses by a few per cent beyond
    (2,048: 312 / 307; 4,096: 625 / 604)."""
    n_sm = torch.cuda.get_device_properties(device).multi_processor_count
    return B <= 8 * n_sm


def n_steps_for(mode, max_turn, first_turn=0):
    """number of move steps after which every board of the batch is finished"""
    last = max_turn + (1 if mode == MODE_MCTS else 2)
    return max(0, last - first_turn)


def run_playouts(pos, policy, mode=MODE_MCTS, max_turn=None, seed=0, game0=0, policy_odd=None, n_steps=None,
                 first_turn=0, komi=5.5, graph=True, persistent=None):
    """Play every board of `pos` to the end with moves drawn from the policy net(s).

    policy:     PackedNet used for every move, or for the moves made at even `turn` when policy_odd is given
                (either may be a reinforce.PolicyTrainer: the net being trained plays in train() mode, selfplay.py:148-150)
    policy_odd: PackedNet for the moves at odd turn (self-play of two nets; needs every board at the same turn parity,
                `first_turn` states it)
    mode:       MODE_MCTS (Go_MCTS.find_random_child, mcts.py:319-364) or MODE_SELFPLAY (legal_sample, selfplay.py:35-47)
    persistent: all moves in one launch of the persistent playout kernel (ignored when a net is being trained);
                otherwise two launches per move, replayed from a CUDA graph when `graph`; None = whichever is faster at this
                batch size (persistent_pays)
    Updates `pos` in place (pos.libs is allocated when absent: the first encode then takes exact liberties, like a fresh
    Game) and returns a PlayoutResult.  Stream-ordered; does not synchronise.
    """
    dev, B = pos.device, pos.B
    if max_turn is None:
        max_turn = MCTS_MAX_TURN if mode == MODE_MCTS else SELFPLAY_MAX_TURN
    if n_steps is None:
        n_steps = n_steps_for(mode, max_turn, first_turn)
    L = _lib.lib()
    fresh_first = pos.libs is None
    if fresh_first:
        pos.libs = torch.zeros(B, 81, dtype=torch.uint8, device=dev)
    moves = torch.empty(n_steps, B, dtype=torch.int16, device=dev)
    bufs = {"conv": torch.empty(L.bk_feats_conv_bytes(B), dtype=torch.uint8, device=dev), "libs": pos.libs}
    probs = torch.empty(B, 81, dtype=torch.float32, device=dev)

    training = any(hasattr(n, "play_probs") for n in (policy, policy_odd) if n is not None)
    if persistent is None:
        persistent = persistent_pays(B, dev)
    if persistent and not training and n_steps > 0:
        playout_run(pos, policy, n_steps, mode, max_turn, seed=seed, game0=game0, policy_odd=policy_odd, first_turn=first_turn,
                    fresh_libs=fresh_first, moves_out=moves)
        score, reward = score_batch(pos.boards, komi)
        return PlayoutResult(moves.t().contiguous(), pos.turn.clone(), score, reward)
    encoded = [False]      # bufs["conv"] holds the planes of the current positions (written by the previous move's launch)

    def step(k, fresh):
        net = policy if (policy_odd is None or (first_turn + k) % 2 == 0) else policy_odd
        if hasattr(net, "play_probs"):     # a net being trained (reinforce.PolicyTrainer): train-mode forward, positions recorded
            net.play_probs(pos, fresh, bufs, probs)
            encoded[0] = False
        else:
            if not encoded[0]:
                features_batch(pos, fresh_libs=fresh, want=("conv", "libs"), out=bufs)
            policy_value_batch(bufs["conv"], B, net, None, want_logits=False, probs_out=probs)
        # with a net in training in the loop the trainer encodes for itself (it needs the byte planes), so only fuse otherwise
        fuse = not training
        playout_step(pos, probs, mode, max_turn, seed=seed, game0=game0, moves_out=moves[k],
                     encode_into=bufs["conv"] if fuse else None)
        encoded[0] = fuse

    k0 = 0
    if fresh_first and n_steps > 0:
        step(0, True)
        k0 = 1
    if graph and n_steps - k0 > 0:
        # warm-up launch outside capture is not needed: the library sets its kernel attributes on first use above or here
        if k0 == 0:
            step(0, False)
            k0 = 1
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap):
            with torch.cuda.graph(g, stream=cap):
                for k in range(k0, n_steps):
                    step(k, False)
        torch.cuda.current_stream(dev).wait_stream(cap)
        g.replay()
    else:
        for k in range(k0, n_steps):
            step(k, False)
    score, reward = score_batch(pos.boards, komi)
    return PlayoutResult(moves.t().contiguous(), pos.turn.clone(), score, reward)


class PlayoutGraph:
    """Games from the empty board, captured once as a CUDA graph (state reset, every move step, scoring) and replayable:
    the form used when the same batch of games is played repeatedly (benchmarks, fixed-size self-play workers).
    Seed and first game id are kernel arguments and therefore fixed at capture."""

    def __init__(self, B, device, policy, mode, max_turn=None, seed=0, game0=0, policy_odd=None, komi=5.5, persistent=None):
        dev = _lib.require_device(device)
        if persistent is None:
            persistent = persistent_pays(B, dev)
        self.persistent = persistent
        if max_turn is None:
            max_turn = MCTS_MAX_TURN if mode == MODE_MCTS else SELFPLAY_MAX_TURN
        self.pos = Positions.empty(B, dev)
        self.n_steps = n_steps_for(mode, max_turn, 0)
        L = _lib.lib()
        pos = self.pos
        self.moves = torch.empty(self.n_steps, B, dtype=torch.int16, device=dev)
        bufs = {"conv": torch.empty(L.bk_feats_conv_bytes(B), dtype=torch.uint8, device=dev), "libs": pos.libs}
        probs = torch.empty(B, 81, dtype=torch.float32, device=dev)
        self.score = torch.empty(B, dtype=torch.float32, device=dev)
        self.reward = torch.empty(B, dtype=torch.int8, device=dev)
        self.rec = torch.empty(B, self.n_steps + 3, dtype=torch.int16, device=dev)

        def body():
            pos.boards.zero_(); pos.ko.fill_(-1); pos.last.fill_(-2); pos.turn.zero_(); pos.done.zero_()
            if persistent:
                playout_run(pos, policy, self.n_steps, mode, max_turn, seed=seed, game0=game0, policy_odd=policy_odd, fresh_libs=True,
                            moves_out=self.moves)
            else:
                features_batch(pos, fresh_libs=True, want=("conv", "libs"), out=bufs)
                for k in range(self.n_steps):
                    net = policy if (policy_odd is None or k % 2 == 0) else policy_odd
                    policy_value_batch(bufs["conv"], B, net, None, want_logits=False, probs_out=probs)
                    playout_step(pos, probs, mode, max_turn, seed=seed, game0=game0, moves_out=self.moves[k], encode_into=bufs["conv"])
            score_batch(pos.boards, komi, out=(self.score, self.reward))
            pack_records(self.moves, pos.turn, self.score, self.reward, out=self.rec)

        # one eager step first: the library sets its kernel attributes on first use, which must not happen under capture
        features_batch(pos, fresh_libs=True, want=("conv", "libs"), out=bufs)
        policy_value_batch(bufs["conv"], B, policy, None, want_logits=False, probs_out=probs)
        self.launches = 5 + (1 if persistent else 1 + 2 * self.n_steps) + 2
        self.graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap):
            with torch.cuda.graph(self.graph, stream=cap):
                body()
        torch.cuda.current_stream(dev).wait_stream(cap)

    def replay(self):
        self.graph.replay()
        return PlayoutResult(self.moves.t(), self.pos.turn, self.score, self.reward, records=self.rec)


def self_play(n_games, policy_black, policy_white, device, seed=0, rank=0, world=1, graph=True):
    """bin/selfplay.py:49-57 `self_play` for this rank's share of `n_games` games from the empty board; the result of a game
    is the sign of Game.score() (gnugo is not available, SURVEY 8c shim 4).  Returns (lo, hi, PlayoutResult)."""
    lo, hi = shard_range(n_games, rank, world)
    pos = Positions.empty(hi - lo, device, track_libs=False)
    res = run_playouts(pos, policy_black, MODE_SELFPLAY, SELFPLAY_MAX_TURN, seed=seed, game0=lo, policy_odd=policy_white,
                       graph=graph)
    return lo, hi, res


def simulate(n_boards, policy, device, seed=0, rank=0, world=1, graph=True):
    """`--simulate` playouts (mcts.py:195-206) from the empty board for this rank's share of `n_boards` boards"""
    lo, hi = shard_range(n_boards, rank, world)
    pos = Positions.empty(hi - lo, device, track_libs=False)
    res = run_playouts(pos, policy, MODE_MCTS, MCTS_MAX_TURN, seed=seed, game0=lo, graph=graph)
    return lo, hi, res


def gather_records(local, n_total, rank, world, group=None):
    """all_gather of the ranks' record blocks (int16 [n_local, W], block r = shard_range(n_total, r, world)) into the full
    [n_total, W] tensor on every rank.  NCCL for CUDA tensors, gloo for CPU tensors; a no-op for world == 1."""
    if world == 1:
        return local
    import torch.distributed as dist
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = local.shape[1]
    if local.is_cuda and n_total % world == 0 and local.is_contiguous():
        # equal shards on GPUs: ONE collective straight into the result (NCCL has no int16: the bytes travel as uint8)
        out = torch.empty(n_total, width, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out.view(torch.uint8), local.view(torch.uint8), group=group)
        return out
    most = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(most, width, dtype=torch.int32, device=local.device)   # int32 on the wire: gloo has no int16
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0).to(local.dtype)
