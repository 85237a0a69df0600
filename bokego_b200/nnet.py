"""Host-side mirror of the reference's `bokego.nnet` module surface (/root/reference/bokego/nnet.py).

Same class / function names, signatures, state-dict key names and return types, so that mcts.py, gtp.py,
boke.py and bin/selfplay.py can import this module in place of `bokego.nnet`.  Everything that computes
runs in the CUDA kernels behind the C ABI (include/bokego_b200.h):

    features(game)                        -> bk_encode        (B = 1)
    PolicyNet.forward / ValueNet.forward  -> bk_repack_f32 + bk_forward
    policy_dist / value / policy_sample   -> thin wrappers, as in the reference

There is no CPU path: a `device` that is not an sm_100 CUDA device raises BokegoB200Error.  forward() in training mode
evaluates ONE position per call with that position's BatchNorm statistics (how the reference's REINFORCE loop calls a net in
train() mode); batched training lives in bokego_b200.reinforce, and no autograd graph is ever built.
The batched entry points that carry the throughput are re-exported from bokego_b200.batched.
"""
from math import sqrt

import numpy as np
import torch
import torch.nn as nn
from torch.distributions.categorical import Categorical
from torch.nn.modules.utils import _pair
from torch.nn.parameter import Parameter

from . import _lib
from . import go
from .batched import (PackedNet, Positions, features_batch, playout_step, policy_value_batch, repack_planes,  # noqa: F401
                      score_batch)

SOFT = nn.Softmax(dim=1)


def _default_device():
    if not torch.cuda.is_available():
        raise _lib.BokegoB200Error("CUDA is not available; bokego_b200 has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


class Conv2dUntiedBias(nn.Module):
    """1x1-style convolution with one bias per output channel AND board square (nnet.py:138-180).
    Holds the parameters `weight (out, in/groups, k, k)` and `bias (out, height, width)`; inside PolicyNet /
    ValueNet it is evaluated as the fused epilogue of the last conv layer."""

    def __init__(self, height, width, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1):
        super(Conv2dUntiedBias, self).__init__()
        if in_channels % groups != 0:
            raise ValueError('in_channels must be divisible by groups')
        if out_channels % groups != 0:
            raise ValueError('out_channels must be divisible by groups')
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = _pair(kernel_size), _pair(stride)
        self.padding, self.dilation, self.groups = _pair(padding), _pair(dilation), groups
        self.weight = Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        self.bias = Parameter(torch.Tensor(out_channels, height, width))
        self.reset_parameters()

    def reset_parameters(self):
        fan_in = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
        bound = 1. / sqrt(fan_in)
        self.weight.data.uniform_(-bound, bound)
        self.bias.data.uniform_(-bound, bound)

    def forward(self, input):
        raise _lib.BokegoB200Error("Conv2dUntiedBias is evaluated inside the fused PolicyNet/ValueNet kernel; "
                                   "call the enclosing net")


def _trunk():
    """the reference's `conv` Sequential (nnet.py:31-53): same module indices => same state-dict keys"""
    layers = [nn.Conv2d(27, 128, 5, padding=2), nn.BatchNorm2d(128), nn.ReLU()]
    for _ in range(6):
        layers += [nn.Conv2d(128, 128, 3, padding=1), nn.BatchNorm2d(128), nn.ReLU()]
    layers.append(Conv2dUntiedBias(9, 9, 128, 1, 1))
    return nn.Sequential(*layers)


class _FusedNet(nn.Module):
    """common part: lazily packed device blob, refreshed when parameters change or move"""

    _is_value = False

    def _packed(self, dev):
        stamp = (str(dev),) + tuple(int(t._version) for t in self.state_dict(keep_vars=True).values()) + \
            tuple(t.data_ptr() for t in self.parameters())
        cache = self.__dict__.get("_bk_cache")
        if cache is None or cache[0] != stamp:
            cache = (stamp, PackedNet(self.state_dict(), dev, is_value=self._is_value))
            self.__dict__["_bk_cache"] = cache
        return cache[1]

    def _run_train_mode(self, x):
        '''forward in train() mode for ONE position, as the reference's policy_dist / policy_sample call a net that is being
        trained (nnet.py:265-297 with bin/selfplay.py:148-150): BatchNorm normalises with the statistics of this position and
        filters them into its running averages (momentum 0.1, unbiased variance), exactly as torch does per call.  No autograd
        graph is built: gradients come from bokego_b200.reinforce.'''
        from . import reinforce as rf
        if self._is_value:
            raise _lib.BokegoB200Error("ValueNet has no train-mode kernel: call .eval() first")
        x = x.reshape(-1, 27, 9, 9)
        if x.shape[0] != 1:
            raise _lib.BokegoB200Error("train-mode forward is defined for one position per call (how the reference calls it); "
                                       "use bokego_b200.reinforce for batched training, or .eval() for batched inference")
        dev = _lib.require_device(x.device)
        stamp = (str(dev),) + tuple(int(p._version) for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())
        cache = self.__dict__.get("_bk_train_cache")
        if cache is None or cache[0] != stamp:
            cache = (stamp, rf.PolicyTrainer(self.state_dict(), dev))
            self.__dict__["_bk_train_cache"] = cache
        tr = cache[1]
        sd = self.state_dict()
        r, _ = rf.running_from_state_dict(sd)                  # the module's buffers are the truth (they may have been loaded)
        tr.running.copy_(torch.from_numpy(r))
        logits, _, stats = tr.forward(x.to(torch.uint8).reshape(1, 27, 81).contiguous(), rf.BN_POSITION, want_stats=True)
        tr.update_running(stats)
        with torch.no_grad():
            for l, i in enumerate(rf.CONV_IDX):
                sd[f"conv.{i + 1}.running_mean"].copy_(tr.running[0, l])
                sd[f"conv.{i + 1}.running_var"].copy_(tr.running[1, l])
                sd[f"conv.{i + 1}.num_batches_tracked"].add_(1)
        return logits

    def _run(self, x):
        if self.training:
            return self._run_train_mode(x)
        dev = _lib.require_device(x.device)
        x = x.reshape(-1, 27, 9, 9).contiguous().float()
        conv = repack_planes(x)
        net = self._packed(dev)
        if self._is_value:
            _, _, v = policy_value_batch(conv, x.shape[0], None, net)
            return v.reshape(-1, 1)
        logits, _, _ = policy_value_batch(conv, x.shape[0], net, None)
        return logits


class PolicyNet(_FusedNet):
    '''(27,9,9) features --> 81 logits; softmax of the output is the prior over moves (nnet.py:19-57).
    1 5x5 conv, 6 3x3 convs (128 ch, BatchNorm, ReLU), 1 1x1 conv with untied bias.'''

    def __init__(self):
        super(PolicyNet, self).__init__()
        self.conv = _trunk()

    def forward(self, x):
        return self._run(x)


class ValueNet(_FusedNet):
    '''(27,9,9) features --> value in (-1,1) from the current player's perspective (nnet.py:59-113).'''

    _is_value = True

    def __init__(self):
        super(ValueNet, self).__init__()
        self.conv = _trunk()
        self.lin1 = nn.Linear(81, 64)
        self.lin2 = nn.Linear(64, 1)
        self.bn = nn.BatchNorm2d(1)
        self.lin_bn = nn.BatchNorm1d(64)
        self.relu = nn.ReLU()
        self.tanh = nn.Tanh()

    def load_policy_dict(self, policy_dict):
        '''load convolution weights from a PolicyNet state dict'''
        merged = self.state_dict()
        merged.update(policy_dict)
        self.load_state_dict(merged)

    def forward(self, x):
        return self._run(x)


class PolicyNet_v2(nn.Module):
    """v0.2 policy net of the reference (nnet.py:116-136).  No weights ship for it and nothing on the hot path
    uses it; the class exists so that `from bokego.nnet import PolicyNet_v2` (boke.py:7) keeps working."""

    def __init__(self):
        super(PolicyNet_v2, self).__init__()
        layers = [nn.Conv2d(27, 64, 5, padding=2), nn.ReLU(), nn.Conv2d(64, 128, 3, padding=1), nn.ReLU()]
        for _ in range(4):
            layers += [nn.Conv2d(128, 128, 3, padding=1), nn.ReLU()]
        layers.append(Conv2dUntiedBias(9, 9, 128, 1, 1))
        self.conv = nn.Sequential(*layers)

    def forward(self, x):
        raise _lib.BokegoB200Error("PolicyNet_v2 has no B200 kernel (out of scope: no weights ship for it)")


def features(game: go.Game):
    '''go.Game --> (27,9,9) float32 CPU tensor, the reference's planes (nnet.py:182-262):
    0 player stones, 1 opponent stones, 2 empty, 3 black to move, 4 last move, 5 legal,
    6-12 liberties, 13-19 liberties after playing, 20-26 number of captures.
    Like the reference it refreshes `game._libs` (the lazy liberty cache) as a side effect.'''
    dev = _default_device()
    lut = {go.BLACK: 1, go.WHITE: -1, go.EMPTY: 0}
    bd = np.array([[lut[c] for c in game.board]], dtype=np.int8)
    ko = -1 if game.ko is None else int(game.ko)
    last = go.PASS if game.last_move == go.PASS else (-2 if not isinstance(game.last_move, int) else int(game.last_move))
    libs = None if game._libs is None else np.frombuffer(bytes(game._libs), dtype=np.uint8).reshape(1, 81)
    pos = Positions.from_numpy(bd, [ko], [last], [int(game.turn)], dev, libs)
    out = features_batch(pos, want=("f32", "libs"))
    game._libs = bytearray(out["libs"][0].cpu().numpy().tobytes())
    return out["f32"][0].cpu()


def policy_dist(policy: PolicyNet, game: go.Game, device=None, fts: torch.Tensor = None):
    '''torch Categorical over the 81 coordinates (softmax of the policy logits, not masked)'''
    device = _lib.require_device(device if device is not None else _default_device())
    if fts is None:
        fts = features(game)
    probs = SOFT(policy(fts.unsqueeze(0).to(device))).squeeze(0)
    return Categorical(probs)


def value(v: ValueNet, game: go.Game, device=None, fts: torch.Tensor = None):
    '''value net output for the position as a Python float'''
    device = _lib.require_device(device if device is not None else _default_device())
    if fts is None:
        fts = features(game)
    return v(fts.unsqueeze(0).to(device)).item()


def policy_sample(policy: PolicyNet, game: go.Game, device=None, fts: torch.Tensor = None):
    '''one move sampled from the (unmasked) policy distribution, as a 0-d long tensor on `device`'''
    return policy_dist(policy, game, device, fts).sample()


def prefill_caches(tree_cls, nodes, policy_net, value_net=None, device=None):
    '''Batched leaf evaluation for the reference's tree search: evaluates `nodes` (Go_MCTS objects) in one
    encoder launch and one policy+value launch and stores the results in the class-level caches
    `tree_cls._fts_cache/_dist_cache/_val_cache` (mcts.py:42-44), which the unchanged search code then hits
    (mcts.py:371-403) instead of evaluating one position at a time.  Returns the number of nodes
    evaluated.  Nodes already cached are skipped.'''
    dev = _lib.require_device(device if device is not None else _default_device())
    todo = [n for n in nodes if n not in tree_cls._fts_cache or n not in tree_cls._dist_cache
            or (value_net is not None and n not in tree_cls._val_cache)]
    if not todo:
        return 0
    lut = {go.BLACK: 1, go.WHITE: -1, go.EMPTY: 0}
    # the kernel takes one liberty-cache mode per launch: fresh objects first, carried caches second
    order = sorted(range(len(todo)), key=lambda i: todo[i]._libs is not None)
    n_fresh = sum(1 for n in todo if n._libs is None)
    pnet = policy_net._packed(dev)
    vnet = value_net._packed(dev) if value_net is not None else None
    for lo, hi, carried in ((0, n_fresh, False), (n_fresh, len(todo), True)):
        part = [todo[i] for i in order[lo:hi]]
        if not part:
            continue
        bd = np.array([[lut[c] for c in n.board] for n in part], dtype=np.int8)
        ko = [-1 if n.ko is None else int(n.ko) for n in part]
        last = [go.PASS if n.last_move == go.PASS else (-2 if not isinstance(n.last_move, int) else int(n.last_move))
                for n in part]
        libs = np.stack([np.frombuffer(bytes(n._libs), dtype=np.uint8) for n in part]) if carried else None
        pos = Positions.from_numpy(bd, ko, last, [int(n.turn) for n in part], dev, libs)
        out = features_batch(pos, want=("conv", "f32", "libs"))
        logits, _, vals = policy_value_batch(out["conv"], len(part), pnet, vnet, want_logits=True)
        # the distribution exactly as policy_dist builds it (nnet.py:265-275: SOFT of the logits, then Categorical's
        # renormalisation, both on the device) and as Go_MCTS.dist stores it (probabilities moved to the CPU, mcts.py:381)
        probs = SOFT(logits)
        probs = (probs / probs.sum(-1, keepdim=True)).cpu()
        f32, libs_out = out["f32"].cpu(), out["libs"].cpu().numpy()
        vals = None if vals is None else vals.cpu()
        for i, n in enumerate(part):
            n._libs = bytearray(libs_out[i].tobytes())
            tree_cls._fts_cache[n] = f32[i]
            dist = Categorical(probs[i])
            dist.probs = probs[i]
            tree_cls._dist_cache[n] = dist
            if vals is not None:
                tree_cls._val_cache[n] = vals[i].item()
    return len(todo)
