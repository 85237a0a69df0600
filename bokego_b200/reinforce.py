"""REINFORCE self-play training on the device: host mirror of `reinforce` in the reference's bin/selfplay.py:59-122
 over the training kernels of csrc/bk_train.cu.

    reinforce(pi, pi_opp, optimizer, train_color, **kwargs)      same signature and keywords as the reference
    PolicyTrainer                                                flat device buffers + the C-ABI calls

What one iteration does, as in the reference: `bs` self-play games between `pi` (train() mode: BatchNorm normalises every
position with its own statistics, because the reference evaluates one position per call) and `pi_opp` (eval() mode, the fused
inference kernel), all games in lock step on the device; the result of a game is the sign of Game.score() (gnugo is absent,
SURVEY 8c shim 4); then loss = sum over the training colour's moves of -log_prob(move) * reward / bs, its gradient through the
net, and one AdamW step.  The reference resets `loss` for every game (selfplay.py:86), so only the LAST game of a batch reaches
`backward()`: accumulate="reference" (default) reproduces that, accumulate="batch" sums over all games as the comment of
selfplay.py:60 intends.  BatchNorm running statistics are filtered over the train-mode calls in the reference's call order
(self-play games one after the other, then the replay), except that the second call the reference makes on a position whose
first sample was illegal (selfplay.py:39-41) is not repeated: after the 36 * bs replay calls the total weight of all self-play
calls in the filter is 0.9 ** (36 * bs).

There is no CPU path and no autograd: gradients come from the hand-written backward kernels.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

CONV_IDX = (0, 3, 6, 9, 12, 15, 18)     # Conv2d modules inside PolicyNet.conv; BatchNorm2d at +1 (nnet.py:31-52)
HEAD_IDX = 21
TP_W0, TP_W1, TP_VEC, TP_HEADW, TP_HEADB, TP_COUNT = 0, 102400, 987136, 989824, 989952, 990033   # include/bokego_b200.h
BN_POSITION, BN_EVAL = 0, 1
PREC_TF32, PREC_3XTF32, PREC_FFMA = 0, 1, 2          # warp-level mma.sync kernels / FFMA validation path
PREC_TC_TF32, PREC_TC_3XTF32 = 4, 5                    # the same GEMMs on tcgen05 (TMEM accumulators)
MOMENTUM = 0.1


def param_keys():
    """parameter names in `PolicyNet.parameters()` order"""
    keys = []
    for i in CONV_IDX:
        keys += [f"conv.{i}.weight", f"conv.{i}.bias", f"conv.{i + 1}.weight", f"conv.{i + 1}.bias"]
    return keys + [f"conv.{HEAD_IDX}.weight", f"conv.{HEAD_IDX}.bias"]


def _np(v):
    return v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)


def flat_from_tensors(get):
    """{state-dict key -> array} (through `get(key)`) -> flat float32 [TP_COUNT] in the layout of the training GEMMs:
    conv weights as [tap][ci][co] with tap = kh * R + kw (layer 0: ci padded 27 -> 32)."""
    flat = np.zeros(TP_COUNT, np.float32)
    for l, i in enumerate(CONV_IDX):
        w = _np(get(f"conv.{i}.weight")).astype(np.float32)               # [co][ci][R][R]
        k = np.transpose(w, (2, 3, 1, 0)).reshape(-1, w.shape[1], 128)    # [tap][ci][co]
        if l == 0:
            blk = np.zeros((25, 32, 128), np.float32)
            blk[:, :27] = k
            flat[TP_W0:TP_W1] = blk.ravel()
        else:
            o = TP_W1 + (l - 1) * 9 * 128 * 128
            flat[o:o + 9 * 128 * 128] = k.ravel()
        v = TP_VEC + l * 3 * 128
        flat[v:v + 128] = _np(get(f"conv.{i}.bias"))
        flat[v + 128:v + 256] = _np(get(f"conv.{i + 1}.weight"))
        flat[v + 256:v + 384] = _np(get(f"conv.{i + 1}.bias"))
    flat[TP_HEADW:TP_HEADW + 128] = _np(get(f"conv.{HEAD_IDX}.weight")).reshape(128)
    flat[TP_HEADB:TP_HEADB + 81] = _np(get(f"conv.{HEAD_IDX}.bias")).reshape(81)
    return flat


def tensors_from_flat(flat):
    """inverse of flat_from_tensors: {parameter key -> float32 array in the state-dict shape}"""
    flat = _np(flat)
    out = {}
    for l, i in enumerate(CONV_IDX):
        if l == 0:
            k = flat[TP_W0:TP_W1].reshape(25, 32, 128)[:, :27]
            out[f"conv.{i}.weight"] = np.ascontiguousarray(np.transpose(k.reshape(5, 5, 27, 128), (3, 2, 0, 1)))
        else:
            o = TP_W1 + (l - 1) * 9 * 128 * 128
            k = flat[o:o + 9 * 128 * 128].reshape(3, 3, 128, 128)
            out[f"conv.{i}.weight"] = np.ascontiguousarray(np.transpose(k, (3, 2, 0, 1)))
        v = TP_VEC + l * 3 * 128
        out[f"conv.{i}.bias"] = flat[v:v + 128].copy()
        out[f"conv.{i + 1}.weight"] = flat[v + 128:v + 256].copy()
        out[f"conv.{i + 1}.bias"] = flat[v + 256:v + 384].copy()
    out[f"conv.{HEAD_IDX}.weight"] = flat[TP_HEADW:TP_HEADW + 128].reshape(1, 128, 1, 1).copy()
    out[f"conv.{HEAD_IDX}.bias"] = flat[TP_HEADB:TP_HEADB + 81].reshape(1, 9, 9).copy()
    return out


def running_from_state_dict(sd):
    """BatchNorm running statistics -> float32 [2][7][128] (means, then variances) and num_batches_tracked int64 [7]"""
    r = np.zeros((2, 7, 128), np.float32)
    nbt = np.zeros(7, np.int64)
    for l, i in enumerate(CONV_IDX):
        r[0, l] = _np(sd[f"conv.{i + 1}.running_mean"])
        r[1, l] = _np(sd[f"conv.{i + 1}.running_var"])
        if f"conv.{i + 1}.num_batches_tracked" in sd:
            nbt[l] = int(_np(sd[f"conv.{i + 1}.num_batches_tracked"]))
    return r, nbt


def game_major(n_steps, n_games):
    """index list that re-orders lock-step rows [step][game] into the reference's call order [game][step]"""
    return (np.arange(n_steps)[None, :] * n_games + np.arange(n_games)[:, None]).reshape(-1).astype(np.int32)


class PolicyTrainer:
    """Device-resident training state of one PolicyNet: flat parameters, gradients, Adam moments, running statistics."""

    def __init__(self, state_dict, device, prec=PREC_TC_3XTF32):
        dev = _lib.require_device(device)
        self.device, self.prec = dev, prec
        self.params = torch.from_numpy(flat_from_tensors(lambda k: state_dict[k])).to(dev)
        r, nbt = running_from_state_dict(state_dict)
        self.running = torch.from_numpy(r).to(dev)
        self.num_batches_tracked = nbt
        self.grads = torch.zeros(TP_COUNT, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(TP_COUNT, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(TP_COUNT, dtype=torch.float32, device=dev)
        self.step = 0
        self._ws_cache = {}              # P -> workspace (the play step and the training step use different sizes)
        self._ws = None
        self._fwd = None                 # (P, bn_mode) of the forward whose activations the workspace holds
        self._eval_net = None

    # ---- kernels ----
    def _workspace(self, P):
        L = _lib.lib()
        ws = self._ws_cache.get(P)
        if ws is None:
            if len(self._ws_cache) >= 3:
                self._ws_cache.pop(next(iter(self._ws_cache)))
            ws = torch.empty(L.bk_train_workspace_bytes(P), dtype=torch.uint8, device=self.device)
            self._ws_cache[P] = ws
        self._ws = ws
        return ws

    def forward(self, planes_u8, bn_mode=BN_POSITION, want_probs=False, want_stats=False, probs_out=None, stats_out=None):
        """PolicyNet.forward (nnet.py:54-57) for P positions given as uint8 planes [P,27,81] (features_batch "u8").
        Returns (logits [P,81], probs | None, stats [P,7,2,128] | None); activations stay in the workspace."""
        L = _lib.lib()
        dev = self.device
        P = planes_u8.shape[0]
        if planes_u8.dtype != torch.uint8 or tuple(planes_u8.shape) != (P, 27, 81) or not planes_u8.is_contiguous() \
                or planes_u8.device != dev:
            raise ValueError("planes_u8: expected contiguous uint8 [P,27,81] on the trainer's device")
        ws = self._workspace(P)
        logits = torch.empty(P, 81, dtype=torch.float32, device=dev)
        probs = probs_out if probs_out is not None else (torch.empty(P, 81, dtype=torch.float32, device=dev) if want_probs else None)
        stats = stats_out if stats_out is not None else \
            (torch.empty(P, 7, 2, 128, dtype=torch.float32, device=dev) if want_stats else None)
        with torch.cuda.device(dev):
            rc = L.bk_train_forward(_lib.ptr(self.params), _lib.ptr(self.running), _lib.ptr(planes_u8), P, bn_mode, self.prec,
                                    _lib.ptr(ws), _lib.ptr(logits), _lib.ptr(probs), _lib.ptr(stats), _lib.stream_ptr(dev))
        _lib.check(rc, "bk_train_forward")
        _lib.count_launch(L.bk_train_launches(0, P, self.prec))
        self._fwd = (P, bn_mode)
        return logits, probs, stats

    def backward(self, moves, coef, accumulate=False):
        """gradient of sum_p coef[p] * -log_prob_p(moves[p]) into self.grads; returns nlp float32 [P] (= -log_prob)."""
        L = _lib.lib()
        dev = self.device
        if self._fwd is None:
            raise RuntimeError("backward() needs the activations of a forward() call")
        P, bn_mode = self._fwd
        if moves.dtype != torch.int16 or tuple(moves.shape) != (P,) or moves.device != dev or not moves.is_contiguous():
            raise ValueError("moves: expected contiguous int16 [P] on the trainer's device")
        if coef.dtype != torch.float32 or tuple(coef.shape) != (P,) or coef.device != dev or not coef.is_contiguous():
            raise ValueError("coef: expected contiguous float32 [P] on the trainer's device")
        nlp = torch.empty(P, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = L.bk_train_backward(_lib.ptr(self.params), _lib.ptr(moves), _lib.ptr(coef), P, bn_mode, self.prec,
                                     _lib.ptr(self._ws), _lib.ptr(self.grads), 1 if accumulate else 0, _lib.ptr(nlp),
                                     _lib.stream_ptr(dev))
        _lib.check(rc, "bk_train_backward")
        _lib.count_launch(L.bk_train_launches(1, P, self.prec))
        return nlp

    def update_running(self, stats, seq=None, momentum=MOMENTUM):
        """BatchNorm's running-average filter over train-mode calls on rows seq[0], seq[1], ... of `stats`, in that order"""
        L = _lib.lib()
        dev = self.device
        S = stats.shape[0] if seq is None else int(seq.shape[0])
        if seq is not None and (seq.dtype != torch.int32 or seq.device != dev):
            raise ValueError("seq: expected int32 on the trainer's device")
        with torch.cuda.device(dev):
            rc = L.bk_train_running_stats(_lib.ptr(self.running), _lib.ptr(stats), _lib.ptr(seq), S, C.c_float(momentum),
                                          _lib.stream_ptr(dev))
        _lib.check(rc, "bk_train_running_stats")
        _lib.count_launch()
        self.num_batches_tracked = self.num_batches_tracked + S
        self._eval_net = None

    def adamw_step(self, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        """torch.optim.AdamW's update of every parameter with self.grads (selfplay.py:138 defaults)"""
        L = _lib.lib()
        dev = self.device
        self.step += 1
        with torch.cuda.device(dev):
            rc = L.bk_adamw_step(_lib.ptr(self.params), _lib.ptr(self.grads), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                 C.c_size_t(TP_COUNT), C.c_double(lr), C.c_double(betas[0]), C.c_double(betas[1]), C.c_double(eps),
                                 C.c_double(weight_decay), self.step, _lib.stream_ptr(dev))
        _lib.check(rc, "bk_adamw_step")
        _lib.count_launch()
        self._fwd = None
        self._eval_net = None

    # ---- state ----
    def state_dict(self):
        """the PolicyNet state dict (CPU float32 tensors, reference key names and shapes)"""
        out = {k: torch.from_numpy(v) for k, v in tensors_from_flat(self.params).items()}
        r = self.running.cpu().numpy()
        for l, i in enumerate(CONV_IDX):
            out[f"conv.{i + 1}.running_mean"] = torch.from_numpy(r[0, l].copy())
            out[f"conv.{i + 1}.running_var"] = torch.from_numpy(r[1, l].copy())
            out[f"conv.{i + 1}.num_batches_tracked"] = torch.tensor(int(self.num_batches_tracked[l]))
        return out

    def grads_dict(self):
        return {k: torch.from_numpy(v) for k, v in tensors_from_flat(self.grads).items()}

    def eval_net(self):
        """PackedNet of the current parameters for the fused inference kernel (eval-mode BatchNorm)"""
        from .batched import PackedNet
        if self._eval_net is None:
            self._eval_net = PackedNet(self.state_dict(), self.device, is_value=False)
        return self._eval_net

    # ---- the hook run_playouts calls for this net's moves: train-mode probabilities, planes and statistics kept ----
    def begin_recording(self, n_steps, B):
        dev = self.device
        self._rec_planes = torch.empty(n_steps, B, 27, 81, dtype=torch.uint8, device=dev)
        self._rec_stats = torch.empty(n_steps, B, 7, 2, 128, dtype=torch.float32, device=dev)
        self._rec_k = 0

    def play_probs(self, pos, fresh, bufs, probs_out):
        from .batched import features_batch
        k = self._rec_k
        self._rec_k += 1
        out = {"u8": self._rec_planes[k], "libs": bufs["libs"]}
        features_batch(pos, fresh_libs=fresh, want=("u8", "libs"), out=out)
        self.forward(self._rec_planes[k], BN_POSITION, probs_out=probs_out, stats_out=self._rec_stats[k])


def compute_grads(trainer, planes_u8, moves, coef, bn_mode=BN_POSITION, chunk=2048):
    """forward + backward over P positions (in chunks of `chunk`): trainer.grads = d loss / d params, returns the loss (0-d)"""
    P = planes_u8.shape[0]
    loss = torch.zeros((), dtype=torch.float32, device=trainer.device)
    for lo in range(0, P, chunk):
        hi = min(P, lo + chunk)
        trainer.forward(planes_u8[lo:hi], bn_mode)
        nlp = trainer.backward(moves[lo:hi], coef[lo:hi], accumulate=lo > 0)
        loss = loss + (nlp * coef[lo:hi]).sum()
    if P == 0:
        trainer.grads.zero_()
    return loss


def reinforce_step(trainer, planes_u8, moves, coef, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01,
                   bn_mode=BN_POSITION, chunk=2048, group=None):
    """forward + backward over this rank's P positions, the sum of the gradients over the ranks of `group` (one all-reduce of the
    flat gradient buffer: NCCL on the device; a no-op without an initialised process group), and one AdamW step -- every rank
    applies the same update to its replica.  Returns the loss summed over ranks as a 0-d tensor."""
    loss = compute_grads(trainer, planes_u8, moves, coef, bn_mode, chunk)
    if _world(group) > 1:
        import torch.distributed as dist
        dist.all_reduce(trainer.grads, group=group)
        dist.all_reduce(loss, group=group)
    trainer.adamw_step(lr, betas, eps, weight_decay)
    return loss


def _world(group=None):
    import torch.distributed as dist
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _rank(group=None):
    import torch.distributed as dist
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def gather_games(local, n_games, group=None):
    """[n_steps, n_local, ...] blocks of the ranks (rank r holds games shard_range(n_games, r, world)) -> [n_steps, n_games, ...]
    on every rank, games in global order.  One all_gather (NCCL for CUDA tensors, gloo for CPU tensors); identity for one rank."""
    world = _world(group)
    if world == 1:
        return local
    import torch.distributed as dist
    from .playout import shard_range
    sizes = [hi - lo for lo, hi in (shard_range(n_games, r, world) for r in range(world))]
    most = max(sizes)
    if local.dtype == torch.int16:                       # NCCL carries no 16-bit integers: int32 on the wire
        return gather_games(local.to(torch.int32), n_games, group).to(torch.int16)
    pad = torch.zeros((local.shape[0], most) + tuple(local.shape[2:]), dtype=local.dtype, device=local.device)
    pad[:, : local.shape[1]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous(), group=group)
    return torch.cat([parts[r][:, : sizes[r]] for r in range(world)], dim=1).contiguous()


def _adopt_optimizer(trainer, pi, optimizer):
    """hyper-parameters and moments of a torch AdamW over pi.parameters() -> the trainer's flat buffers"""
    g = optimizer.param_groups[0]
    hyper = dict(lr=g["lr"], betas=tuple(g["betas"]), eps=g["eps"], weight_decay=g["weight_decay"])
    named = dict(pi.named_parameters())
    st = optimizer.state
    if all(named[k] in st and "exp_avg" in st[named[k]] for k in param_keys()):
        trainer.exp_avg = torch.from_numpy(flat_from_tensors(lambda k: st[named[k]]["exp_avg"])).to(trainer.device)
        trainer.exp_avg_sq = torch.from_numpy(flat_from_tensors(lambda k: st[named[k]]["exp_avg_sq"])).to(trainer.device)
        trainer.step = int(st[named[param_keys()[0]]]["step"])
    return hyper


def _hand_back(trainer, pi, optimizer):
    """trained parameters, running statistics and Adam moments back into the torch module / optimizer"""
    sd = trainer.state_dict()
    with torch.no_grad():
        own = pi.state_dict()
        for k, v in sd.items():
            own[k].copy_(v.to(own[k].device))
    named = dict(pi.named_parameters())
    m, v = tensors_from_flat(trainer.exp_avg), tensors_from_flat(trainer.exp_avg_sq)
    for k in param_keys():
        p = named[k]
        optimizer.state[p] = {"step": torch.tensor(float(trainer.step)),
                              "exp_avg": torch.from_numpy(m[k]).to(p.device).reshape(p.shape),
                              "exp_avg_sq": torch.from_numpy(v[k]).to(p.device).reshape(p.shape)}


def reinforce(pi, pi_opp, optimizer, train_color, **kwargs):
    '''REINFORCE policy gradient by self-play (bin/selfplay.py:59-122), same arguments:
        pi: training PolicyNet (bokego_b200.nnet.PolicyNet), pi_opp: opponent PolicyNet,
        optimizer: torch.optim.AdamW over pi.parameters(), train_color: "black" or "white"
    kwargs: n_itrs (60), bs (16), device, stats (list the win counts are appended to), id;
    additional: accumulate ("reference" | "batch"), seed (random stream of the games), prec (PREC_*), group (process group).
    With torch.distributed initialised (one process per GPU) the `bs` games of an iteration are sharded over the ranks by global
    game id, the per-call statistics are all-gathered, the gradients all-reduced (NCCL), and every rank applies the same step:
    the replicas stay identical and the games do not depend on the number of ranks.
    The trained weights, running statistics and optimizer state are written back into `pi` / `optimizer`.'''
    from .playout import SELFPLAY_MAX_TURN, run_playouts
    from .batched import MODE_SELFPLAY, PackedNet, Positions
    if train_color not in ("black", "white"):
        raise ValueError("train_color must be black or white")
    n_itrs = kwargs.get("n_itrs", 60)
    bs = kwargs.get("bs", 16)
    device = kwargs.get("device")
    stats = kwargs.get("stats")
    idn = kwargs.get("id", '')
    accumulate = kwargs.get("accumulate", "reference")
    seed = kwargs.get("seed", 0)
    dev = _lib.require_device(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    white = train_color == "white"
    group = kwargs.get("group")
    rank, world = _rank(group), _world(group)
    from .playout import shard_range
    lo, hi = shard_range(bs, rank, world)            # this rank's games (global ids: the random stream is keyed by them)
    nb = hi - lo
    trainer = PolicyTrainer(pi.state_dict(), dev, prec=kwargs.get("prec", PREC_TC_3XTF32))
    hyper = _adopt_optimizer(trainer, pi, optimizer)
    opp = PackedNet(pi_opp.state_dict(), dev, is_value=False)
    n_mine = (SELFPLAY_MAX_TURN + 2) // 2                      # 36 moves of the training colour in a 72-move game
    order = torch.from_numpy(game_major(n_mine, bs)).to(dev)
    winlist = []
    for itr in range(n_itrs):
        pos = Positions.empty(nb, dev, track_libs=False)
        trainer.begin_recording(n_mine, nb)
        if nb > 0:
            res = run_playouts(pos, opp if white else trainer, MODE_SELFPLAY, SELFPLAY_MAX_TURN, seed=seed + itr, game0=lo,
                               policy_odd=trainer if white else opp, graph=False)
            results = res.reward.to(torch.float32)              # +1 black wins, -1 otherwise (sign of Game.score())
            # the moves of the training colour, [step][game] like the recorded planes; a finished game has codes < PASS there
            mine = res.moves[:, (1 if white else 0)::2].t().contiguous()
        else:
            results = torch.zeros(0, dtype=torch.float32, device=dev)
            mine = torch.zeros(n_mine, 0, dtype=torch.int16, device=dev)
        reward = -results if white else results
        played = mine >= 0
        coef = (reward / bs)[None, :].expand(n_mine, nb).clone()
        if accumulate == "reference":
            # `loss` is reset per game (selfplay.py:86): only the last game of the batch is differentiated
            sel = slice(nb - 1, nb) if (hi == bs and nb > 0) else slice(0, 0)
        elif accumulate == "batch":
            sel = slice(0, nb)
        else:
            raise ValueError('accumulate must be "reference" or "batch"')
        coef = torch.where(played, coef, torch.zeros_like(coef))
        # running statistics: the self-play calls game by game, then the replay calls on the same positions (all ranks' games)
        stats_rows = gather_games(trainer._rec_stats, bs, group).reshape(n_mine * bs, 7, 2, 128)
        # only the calls the reference makes: no net call follows the end of a game (legal_sample returns None, selfplay.py:44-45,
        # and the replay is bounded by len(g), selfplay.py:91-100), so rows of finished games are left out of the filter
        # (gathered as bytes: NCCL has no 16-bit integer type)
        played_all = gather_games(played.to(torch.uint8).unsqueeze(-1), bs, group)[..., 0] != 0   # [step][game], all ranks' games
        seq = order[played_all.t().reshape(-1)]
        trainer.update_running(stats_rows, torch.cat([seq, seq]))
        # positions are independent under per-position BatchNorm, so games with a zero coefficient are skipped
        reinforce_step(trainer, trainer._rec_planes[:, sel].reshape(-1, 27, 81).contiguous(),
                       mine[:, sel].reshape(-1).clamp(min=0).contiguous(), coef[:, sel].reshape(-1).contiguous(), group=group,
                       **hyper)
        wins = (reward == 1).sum()
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(wins, group=group)
        winlist.append(int(wins.item()))
        if len(winlist) > 0 and len(winlist) % 10 == 0 and rank == 0:
            avg_win = sum(winlist[-10:]) / (bs * 10)
            print(f"Winrate ({train_color}{idn}): {avg_win:.2f}")
    _hand_back(trainer, pi, optimizer)
    if stats is not None:
        stats.extend(winlist)
    return trainer
