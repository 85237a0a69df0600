"""Batched entry points of the hot path (SURVEY 8b "additive batched API").

Thin host code over the C ABI: argument checking, buffer allocation with torch, stream plumbing.
All tensors live on one sm_100 CUDA device; every call is stream-ordered and does not synchronise.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

NONE, PASS = -2, -1
MODE_MCTS, MODE_SELFPLAY = 0, 1
FLAG_POLICY, FLAG_VALUE, FLAG_SIMT = 1, 2, 4


def _want(t, dtype, shape, name, device=None):
    if not isinstance(t, torch.Tensor) or t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous {dtype} tensor of shape {tuple(shape)}")
    if device is not None and t.device != device:
        raise ValueError(f"{name}: expected device {device}, got {t.device}")


class Positions:
    """A batch of go.Game states on the device (go.py:51-66): board, ko, last_move, turn, _libs."""

    def __init__(self, boards, ko, last, turn, libs=None, done=None):
        dev = _lib.require_device(boards.device)
        B = boards.shape[0]
        _want(boards, torch.int8, (B, 81), "boards", dev)
        for t, n in ((ko, "ko"), (last, "last"), (turn, "turn")):
            _want(t, torch.int16, (B,), n, dev)
        if libs is not None:
            _want(libs, torch.uint8, (B, 81), "libs", dev)
        self.boards, self.ko, self.last, self.turn, self.libs = boards, ko, last, turn, libs
        self.done = torch.zeros(B, dtype=torch.uint8, device=dev) if done is None else done
        self.device, self.B = dev, B

    @classmethod
    def empty(cls, B, device, track_libs=True):
        dev = _lib.require_device(device)
        return cls(torch.zeros(B, 81, dtype=torch.int8, device=dev), torch.full((B,), -1, dtype=torch.int16, device=dev),
                   torch.full((B,), NONE, dtype=torch.int16, device=dev), torch.zeros(B, dtype=torch.int16, device=dev),
                   torch.zeros(B, 81, dtype=torch.uint8, device=dev) if track_libs else None)

    @classmethod
    def from_numpy(cls, boards, ko, last, turn, device, libs=None):
        dev = _lib.require_device(device)
        f = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        return cls(f(boards, np.int8).reshape(-1, 81), f(ko, np.int16), f(last, np.int16), f(turn, np.int16),
                   None if libs is None else f(libs, np.uint8).reshape(-1, 81))


def features_batch(pos, fresh_libs=None, want=("conv", "legal", "libs"), out=None):
    """nnet.features for a batch (kernel a).

    fresh_libs: True -> treat every position as a fresh go.Game (exact liberties);
                False -> apply the lazy update to pos.libs (the carried Game._libs);
                None  -> fresh iff pos.libs is None.
    want: any of "conv" (fp16 operand for policy_value_batch), "f32" ([B,27,9,9] float32, what
          nnet.features returns), "u8" ([B,27,81] uint8), "legal", "libs".
    Returns a dict with the requested tensors.  When the liberty cache is carried and "libs" is
    requested, pos.libs is replaced by the updated cache.
    """
    L = _lib.lib()
    dev, B = pos.device, pos.B
    fresh = pos.libs is None if fresh_libs is None else fresh_libs
    if not fresh and pos.libs is None:
        raise ValueError("fresh_libs=False needs pos.libs")
    out = {} if out is None else out
    if "conv" in want and "conv" not in out:
        out["conv"] = torch.empty(L.bk_feats_conv_bytes(B), dtype=torch.uint8, device=dev)
    if "f32" in want and "f32" not in out:
        out["f32"] = torch.empty(B, 27, 9, 9, dtype=torch.float32, device=dev)
    if "u8" in want and "u8" not in out:
        out["u8"] = torch.empty(B, 27, 81, dtype=torch.uint8, device=dev)
    if "legal" in want and "legal" not in out:
        out["legal"] = torch.empty(B, 81, dtype=torch.uint8, device=dev)
    if "libs" in want and "libs" not in out:
        spare = getattr(pos, "_libs_spare", None)
        out["libs"] = spare if spare is not None else torch.empty(B, 81, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = L.bk_encode(_lib.ptr(pos.boards), _lib.ptr(pos.ko), _lib.ptr(pos.last), _lib.ptr(pos.turn),
                         None if fresh else _lib.ptr(pos.libs), _lib.ptr(out.get("conv")), _lib.ptr(out.get("f32")),
                         _lib.ptr(out.get("u8")), _lib.ptr(out.get("legal")), _lib.ptr(out.get("libs")), B,
                         _lib.stream_ptr(dev))
    _lib.check(rc, "bk_encode")
    _lib.count_launch()
    if "libs" in want:
        # without a caller-supplied buffer the cache is written to a spare tensor and the two are swapped (a fresh position
        # adopts its exact liberties as the cache, exactly like Game._libs after the first get_liberties call); callers that
        # pass out["libs"] = pos.libs (the playout loops) get the in-place update the C ABI allows (libs_out may alias libs_in)
        pos._libs_spare, pos.libs = pos.libs, out["libs"]
    return out


def repack_planes(feats_f32):
    """float32 [B,27,9,9] planes on the device -> the fp16 operand of policy_value_batch"""
    L = _lib.lib()
    dev = _lib.require_device(feats_f32.device)
    B = feats_f32.shape[0]
    _want(feats_f32, torch.float32, (B, 27, 9, 9), "feats_f32", dev)
    conv = torch.empty(L.bk_feats_conv_bytes(B), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = L.bk_repack_f32(_lib.ptr(feats_f32), _lib.ptr(conv), B, _lib.stream_ptr(dev))
    _lib.check(rc, "bk_repack_f32")
    _lib.count_launch()
    return conv


class PackedNet:
    """Device-resident weight blob of one PolicyNet / ValueNet (BatchNorm folded, fp16 conv operands)."""

    def __init__(self, state_dict, device, is_value=None):
        dev = _lib.require_device(device)
        L = _lib.lib()
        sd = {k: (v.detach().cpu().double() if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v)).double())
              for k, v in state_dict.items()}
        self.is_value = ("lin1.weight" in sd) if is_value is None else is_value
        eps = 1e-5
        ws, bs = [], []
        for i in (0, 3, 6, 9, 12, 15, 18):        # Conv2d at conv.i, BatchNorm2d at conv.(i+1)  (nnet.py:31-52)
            w, b = sd[f"conv.{i}.weight"], sd[f"conv.{i}.bias"]
            s = sd[f"conv.{i + 1}.weight"] / torch.sqrt(sd[f"conv.{i + 1}.running_var"] + eps)
            ws.append(w * s[:, None, None, None])
            bs.append((b - sd[f"conv.{i + 1}.running_mean"]) * s + sd[f"conv.{i + 1}.bias"])
        f32 = lambda t: np.ascontiguousarray(t.float().numpy())
        w0, w16 = f32(ws[0]), f32(torch.stack(ws[1:]))
        bias = f32(torch.stack(bs))
        head_w, head_b = f32(sd["conv.21.weight"].reshape(128)), f32(sd["conv.21.bias"].reshape(81))
        vtail = None
        if self.is_value:
            s0 = sd["bn.weight"] / torch.sqrt(sd["bn.running_var"] + eps)
            t0 = sd["bn.bias"] - sd["bn.running_mean"] * s0
            s1 = sd["lin_bn.weight"] / torch.sqrt(sd["lin_bn.running_var"] + eps)
            w1 = sd["lin1.weight"] * s1[:, None]
            b1 = (sd["lin1.bias"] - sd["lin_bn.running_mean"]) * s1 + sd["lin_bn.bias"]
            vtail = f32(torch.cat([s0.reshape(1), t0.reshape(1), sd["lin2.bias"].reshape(1), w1.reshape(-1), b1,
                                   sd["lin2.weight"].reshape(-1)]))
        blob = np.zeros(L.bk_weights_blob_bytes(), np.uint8)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        _lib.check(L.bk_weights_pack(p(w0), p(w16), p(bias), p(head_w), p(head_b), p(vtail), p(blob)), "bk_weights_pack")
        self.blob = torch.from_numpy(blob).to(dev)
        self.device = dev


def policy_value_batch(feats_conv, B, policy=None, value=None, want_logits=True, simt=False, _extra_flags=0, probs_out=None,
                       value_out=None):
    """PolicyNet / ValueNet forward for B positions (kernel b).

    feats_conv: the "conv" output of features_batch.  policy / value: PackedNet or None.
    Returns (logits [B,81] | None, probs [B,81] | None, value [B] | None), float32.
    """
    L = _lib.lib()
    if policy is None and value is None:
        raise ValueError("need a policy net, a value net, or both")
    dev = (policy or value).device
    if feats_conv.device != dev or feats_conv.numel() < L.bk_feats_conv_bytes(B):
        raise ValueError("feats_conv: wrong device or too small for B")
    logits = torch.empty(B, 81, dtype=torch.float32, device=dev) if (policy is not None and want_logits) else None
    probs = None
    if policy is not None:
        if probs_out is not None:
            _want(probs_out, torch.float32, (B, 81), "probs_out", dev)
        probs = probs_out if probs_out is not None else torch.empty(B, 81, dtype=torch.float32, device=dev)
    val = None
    if value is not None:
        if value_out is not None:
            _want(value_out, torch.float32, (B,), "value_out", dev)
        val = value_out if value_out is not None else torch.empty(B, dtype=torch.float32, device=dev)
    flags = (FLAG_POLICY if policy is not None else 0) | (FLAG_VALUE if value is not None else 0) | \
            (FLAG_SIMT if simt else 0) | _extra_flags
    with torch.cuda.device(dev):
        rc = L.bk_forward(_lib.ptr(feats_conv), _lib.ptr(policy.blob if policy else None),
                          _lib.ptr(value.blob if value else None), _lib.ptr(logits), _lib.ptr(probs), _lib.ptr(val),
                          B, flags, _lib.stream_ptr(dev))
    _lib.check(rc, "bk_forward")
    _lib.count_launch()
    return logits, probs, val


def evaluate_positions(pos, policy=None, value=None, fresh_libs=None, want_logits=False, want=(), out=None, probs_out=None,
                       value_out=None):
    """nnet.features + PolicyNet / ValueNet forward for a batch of positions in ONE launch (bk_forward_positions): the conv
    kernel encodes the planes of every item on chip while the tensor pipe works on the item before.
    fresh_libs as in features_batch; want: any of "legal", "libs" (encoder outputs; with "libs" requested pos.libs is replaced by
    the refreshed cache).  Returns (logits | None, probs | None, value | None, dict of the encoder outputs)."""
    L = _lib.lib()
    if policy is None and value is None:
        raise ValueError("need a policy net, a value net, or both")
    dev, B = pos.device, pos.B
    fresh = pos.libs is None if fresh_libs is None else fresh_libs
    if not fresh and pos.libs is None:
        raise ValueError("fresh_libs=False needs pos.libs")
    out = {} if out is None else out
    if "legal" in want and "legal" not in out:
        out["legal"] = torch.empty(B, 81, dtype=torch.uint8, device=dev)
    if "libs" in want and "libs" not in out:
        # never in place: with both nets every board is encoded once per net, by different CTAs (the buffers are swapped instead)
        spare = getattr(pos, "_libs_spare", None)
        out["libs"] = spare if (spare is not None and spare is not pos.libs) else torch.empty(B, 81, dtype=torch.uint8, device=dev)
    logits = torch.empty(B, 81, dtype=torch.float32, device=dev) if (policy is not None and want_logits) else None
    probs = None
    if policy is not None:
        if probs_out is not None:
            _want(probs_out, torch.float32, (B, 81), "probs_out", dev)
        probs = probs_out if probs_out is not None else torch.empty(B, 81, dtype=torch.float32, device=dev)
    val = None
    if value is not None:
        if value_out is not None:
            _want(value_out, torch.float32, (B,), "value_out", dev)
        val = value_out if value_out is not None else torch.empty(B, dtype=torch.float32, device=dev)
    flags = (FLAG_POLICY if policy is not None else 0) | (FLAG_VALUE if value is not None else 0)
    with torch.cuda.device(dev):
        rc = L.bk_forward_positions(_lib.ptr(pos.boards), _lib.ptr(pos.ko), _lib.ptr(pos.last), _lib.ptr(pos.turn),
                                    None if fresh else _lib.ptr(pos.libs), _lib.ptr(policy.blob if policy else None),
                                    _lib.ptr(value.blob if value else None), _lib.ptr(logits), _lib.ptr(probs), _lib.ptr(val),
                                    _lib.ptr(out.get("legal")), _lib.ptr(out.get("libs")), B, flags, _lib.stream_ptr(dev))
    _lib.check(rc, "bk_forward_positions")
    _lib.count_launch()
    if "libs" in want:
        pos._libs_spare, pos.libs = pos.libs, out["libs"]
    return logits, probs, val, out


class HostEvaluator:
    """Policy + value evaluation of positions that live in HOST memory (what a search running on the CPU calls): one packed
    pinned staging buffer each way, so a call is one host-to-device copy (boards, ko, last, turn), the nets, and one
    device-to-host copy (probabilities, values).  Fill `h_boards / h_ko / h_last / h_turn` (views of the pinned input buffer),
    call `run()` (stream-ordered, does not synchronise), read `h_probs / h_value` after a synchronisation.
    depth == 1: a call is ONE kernel (bk_forward_positions: planes computed on chip) between the two copies -- the lowest latency
    per call.  depth > 1: the staging buffers are rotated and the copy-in stream also runs the encoder (bk_encode), whose small CTAs
    fit next to the persistent conv CTAs of the previous call, so that in steady state a call costs one conv kernel (bk_forward on
    planes that are ready) -- measured 1.5 % faster than the one-kernel form, whose on-chip encode of a CTA's first item is
    exposed (`run()` then returns the slot index whose outputs it will fill)."""

    def __init__(self, B, policy, value, device, depth=1):
        dev = _lib.require_device(device)
        L = _lib.lib()
        self.B, self.policy, self.value, self.device, self.depth = B, policy, value, dev, depth
        nb = (B * 81 + 7) // 8 * 8                            # boards, padded so that the int16 arrays behind them are aligned
        n_in, n_out = nb + 3 * 2 * B, 4 * (B * 81 + B)
        self.slots = []
        for _ in range(depth):
            h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
            d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
            h_out = torch.empty(B * 81 + B, dtype=torch.float32).pin_memory()
            d_out = torch.empty(B * 81 + B, dtype=torch.float32, device=dev)
            meta = lambda t, k: t[nb + 2 * B * k: nb + 2 * B * (k + 1)].view(torch.int16)
            views = lambda t: (t[: B * 81].view(torch.int8).view(B, 81), meta(t, 0), meta(t, 1), meta(t, 2))
            hb, hk, hl, ht = views(h_in)
            db, dk, dl, dt = views(d_in)
            self.slots.append({"h_in": h_in, "d_in": d_in, "h_out": h_out, "d_out": d_out,
                               "h": (hb, hk, hl, ht), "pos": Positions(db, dk, dl, dt),
                               "d_probs": d_out[: B * 81].view(B, 81), "d_value": d_out[B * 81:],
                               "h_probs": h_out[: B * 81].view(B, 81), "h_value": h_out[B * 81:],
                               "conv": torch.empty(L.bk_feats_conv_bytes(B), dtype=torch.uint8, device=dev) if depth > 1 else None,
                               "in_done": torch.cuda.Event(), "out_done": torch.cuda.Event(), "computed": torch.cuda.Event()})
        self.k = 0
        self.copy_in = torch.cuda.Stream(device=dev) if depth > 1 else None
        self.copy_out = torch.cuda.Stream(device=dev) if depth > 1 else None
        self.h2d_bytes, self.d2h_bytes = n_in, n_out

    def slot(self, i=None):
        return self.slots[self.k % self.depth if i is None else i]

    def _evaluate(self, s):
        evaluate_positions(s["pos"], self.policy, self.value, fresh_libs=True, probs_out=s["d_probs"], value_out=s["d_value"])

    def run(self):
        s = self.slots[self.k % self.depth]
        i = self.k % self.depth
        self.k += 1
        main = torch.cuda.current_stream(self.device)
        if self.depth == 1:
            s["d_in"].copy_(s["h_in"], non_blocking=True)
            self._evaluate(s)
            s["h_out"].copy_(s["d_out"], non_blocking=True)
            return i
        # rotated buffers: copy-in and copy-out run on their own streams, ordered against the kernels by events
        self.copy_in.wait_event(s["computed"])            # the kernels that last read this slot's inputs / planes are done
        with torch.cuda.stream(self.copy_in):
            s["d_in"].copy_(s["h_in"], non_blocking=True)
            # the encoder runs behind the copy on the same side stream: its small CTAs fit next to the persistent conv
            # CTAs of the previous call, so the feature planes are ready when that call's forward ends
            features_batch(s["pos"], fresh_libs=True, want=("conv",), out={"conv": s["conv"]})
            s["in_done"].record()
        main.wait_event(s["in_done"])
        main.wait_event(s["out_done"])                    # the previous results of this slot have left the device
        policy_value_batch(s["conv"], self.B, self.policy, self.value, want_logits=False, probs_out=s["d_probs"],
                           value_out=s["d_value"])
        s["computed"].record()
        self.copy_out.wait_event(s["computed"])
        with torch.cuda.stream(self.copy_out):
            s["h_out"].copy_(s["d_out"], non_blocking=True)
            s["out_done"].record()
        return i

    def drain(self):
        """make the current stream wait for every outstanding copy"""
        if self.depth > 1:
            main = torch.cuda.current_stream(self.device)
            main.wait_stream(self.copy_in)
            main.wait_stream(self.copy_out)


def playout_step(pos, probs, mode, max_turn, seed=0, game0=0, q_inj=None, moves_out=None, encode_into=None):
    """One playout move for every unfinished board (kernel c); updates `pos` in place, returns moves int16 [B].
    encode_into: the "conv" buffer of a previous features_batch of these boards -- the same launch then also encodes the
    position after the move into it (carried liberty cache pos.libs, updated in place), so the next policy evaluation needs
    no encoder launch."""
    L = _lib.lib()
    dev, B = pos.device, pos.B
    _want(probs, torch.float32, (B, 81), "probs", dev)
    qv = 0
    if q_inj is not None:
        if q_inj.dtype != torch.float32 or q_inj.dim() != 3 or q_inj.shape[0] != B or q_inj.shape[2] != 81 \
                or not q_inj.is_contiguous() or q_inj.device != dev:
            raise ValueError("q_inj: expected contiguous float32 [B, n, 81] on the same device")
        qv = q_inj.shape[1]
    if moves_out is None:
        moves_out = torch.empty(B, dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        if encode_into is None:
            rc = L.bk_playout_step(_lib.ptr(pos.boards), _lib.ptr(pos.ko), _lib.ptr(pos.last), _lib.ptr(pos.turn),
                                   _lib.ptr(pos.libs), _lib.ptr(pos.done), _lib.ptr(probs), _lib.ptr(q_inj), qv,
                                   C.c_uint64(seed), C.c_uint32(game0), mode, max_turn, _lib.ptr(moves_out), B,
                                   _lib.stream_ptr(dev))
        else:
            if pos.libs is None:
                raise ValueError("encode_into needs the carried liberty cache pos.libs")
            if encode_into.device != dev or encode_into.numel() < L.bk_feats_conv_bytes(B):
                raise ValueError("encode_into: wrong device or too small for B")
            rc = L.bk_playout_step_encode(_lib.ptr(pos.boards), _lib.ptr(pos.ko), _lib.ptr(pos.last), _lib.ptr(pos.turn),
                                          _lib.ptr(pos.libs), _lib.ptr(pos.done), _lib.ptr(probs), _lib.ptr(q_inj), qv,
                                          C.c_uint64(seed), C.c_uint32(game0), mode, max_turn, _lib.ptr(moves_out),
                                          _lib.ptr(encode_into), B, _lib.stream_ptr(dev))
    _lib.check(rc, "bk_playout_step")
    _lib.count_launch()
    return moves_out


def playout_run(pos, policy, n_steps, mode, max_turn, seed=0, game0=0, policy_odd=None, first_turn=0, fresh_libs=None,
                moves_out=None):
    """n_steps playout moves for every board in ONE kernel launch (bk_playout_run): the conv kernel keeps each item of <= 5
    boards on its SM for the whole playout, positions resident on chip.  fresh_libs: True -> the positions are fresh Games
    (exact liberties); False -> pos.libs is their carried cache; None -> fresh iff pos.libs is None.
    Updates `pos` in place (pos.libs is allocated when absent); returns moves int16 [n_steps, B]."""
    L = _lib.lib()
    dev, B = pos.device, pos.B
    fresh = pos.libs is None if fresh_libs is None else fresh_libs
    if pos.libs is None:
        if not fresh:
            raise ValueError("fresh_libs=False needs pos.libs")
        pos.libs = torch.zeros(B, 81, dtype=torch.uint8, device=dev)
    if moves_out is None:
        moves_out = torch.empty(n_steps, B, dtype=torch.int16, device=dev)
    _want(moves_out, torch.int16, (n_steps, B), "moves_out", dev)
    with torch.cuda.device(dev):
        rc = L.bk_playout_run(_lib.ptr(pos.boards), _lib.ptr(pos.ko), _lib.ptr(pos.last), _lib.ptr(pos.turn), _lib.ptr(pos.libs),
                              _lib.ptr(pos.done), _lib.ptr(policy.blob), _lib.ptr(policy_odd.blob if policy_odd is not None else None),
                              C.c_uint64(seed), C.c_uint32(game0), mode, max_turn, first_turn, n_steps, int(bool(fresh)),
                              _lib.ptr(moves_out), B, _lib.stream_ptr(dev))
    _lib.check(rc, "bk_playout_run")
    _lib.count_launch()
    return moves_out


def make_moves(pos, parent_idx, moves):
    """Go_MCTS.make_move for a batch (mcts.py:340-346): child c = position parent_idx[c] of `pos` with moves[c] played.
    parent_idx int32 [C], moves int16 [C] (device tensors).  Returns (Positions of the C children, status uint8 [C]);
    the children carry the liberty cache exactly as deepcopy + play_move does."""
    L = _lib.lib()
    dev = pos.device
    C_ = parent_idx.shape[0]
    _want(parent_idx, torch.int32, (C_,), "parent_idx", dev)
    _want(moves, torch.int16, (C_,), "moves", dev)
    child = Positions(torch.empty(C_, 81, dtype=torch.int8, device=dev), torch.empty(C_, dtype=torch.int16, device=dev),
                      torch.empty(C_, dtype=torch.int16, device=dev), torch.empty(C_, dtype=torch.int16, device=dev),
                      torch.empty(C_, 81, dtype=torch.uint8, device=dev))
    status = torch.empty(C_, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = L.bk_make_moves(_lib.ptr(pos.boards), _lib.ptr(pos.ko), _lib.ptr(pos.last), _lib.ptr(pos.turn), _lib.ptr(pos.libs),
                             _lib.ptr(parent_idx), _lib.ptr(moves), _lib.ptr(child.boards), _lib.ptr(child.ko),
                             _lib.ptr(child.last), _lib.ptr(child.turn), _lib.ptr(child.libs), _lib.ptr(status), C_,
                             _lib.stream_ptr(dev))
    _lib.check(rc, "bk_make_moves")
    _lib.count_launch()
    return child, status


def score_batch(boards, komi=5.5, out=None):
    """Game.score() and the +-1 reward for a batch: returns (score float32 [B], reward int8 [B])."""
    L = _lib.lib()
    dev = _lib.require_device(boards.device)
    B = boards.shape[0]
    _want(boards, torch.int8, (B, 81), "boards", dev)
    if out is not None:
        score, reward = out
        _want(score, torch.float32, (B,), "score", dev)
        _want(reward, torch.int8, (B,), "reward", dev)
    else:
        score = torch.empty(B, dtype=torch.float32, device=dev)
        reward = torch.empty(B, dtype=torch.int8, device=dev)
    with torch.cuda.device(dev):
        rc = L.bk_score(_lib.ptr(boards), C.c_float(komi), _lib.ptr(score), _lib.ptr(reward), B, _lib.stream_ptr(dev))
    _lib.check(rc, "bk_score")
    _lib.count_launch()
    return score, reward


def pack_records(moves_tb, turn, score, reward, out=None):
    """per-game records int16 [B, T + 3] = (turn reached, reward, 2 * score, moves...) from the move log int16 [T, B] of a playout
    loop, in one launch (the rows of the one result gather of a multi-GPU run)"""
    L = _lib.lib()
    dev = _lib.require_device(moves_tb.device)
    T, B = moves_tb.shape
    _want(moves_tb, torch.int16, (T, B), "moves_tb", dev)
    _want(turn, torch.int16, (B,), "turn", dev)
    _want(score, torch.float32, (B,), "score", dev)
    _want(reward, torch.int8, (B,), "reward", dev)
    if out is None:
        out = torch.empty(B, T + 3, dtype=torch.int16, device=dev)
    _want(out, torch.int16, (B, T + 3), "out", dev)
    with torch.cuda.device(dev):
        rc = L.bk_pack_records(_lib.ptr(moves_tb), _lib.ptr(turn), _lib.ptr(score), _lib.ptr(reward), _lib.ptr(out), T, B,
                               _lib.stream_ptr(dev))
    _lib.check(rc, "bk_pack_records")
    _lib.count_launch()
    return out


def exp_draws(seed, game0, move, tr, B, device):
    """The counter-based Exp(1) stream as the kernels see it: float32 [B,81]."""
    L = _lib.lib()
    dev = _lib.require_device(device)
    q = torch.empty(B, 81, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = L.bk_exp_draws(C.c_uint64(seed), C.c_uint32(game0), C.c_uint32(move), C.c_uint32(tr), _lib.ptr(q), B,
                            _lib.stream_ptr(dev))
    _lib.check(rc, "bk_exp_draws")
    _lib.count_launch()
    return q
