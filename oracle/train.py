"""fp32 CPU restatement of the reference's REINFORCE step.  TEST INFRASTRUCTURE ONLY (the checker, never the product).

Restates, as functions of a state dict,
  * PolicyNet.forward in train() mode called with ONE position per call, as `policy_dist` does
    (/root/reference/bokego/nnet.py:19-57, 265-275; bin/selfplay.py:148-150 puts `pi` in train mode):
    every BatchNorm2d normalises with the statistics of that single position (81 squares per channel,
    biased variance, eps 1e-5) and filters its running statistics with momentum 0.1 and the unbiased
    variance -- per call, in call order;
  * the loss of `reinforce` (bin/selfplay.py:84-117): sum over the training colour's moves of
    -Categorical(softmax(logits)).log_prob(move), times reward / bs.  The reference resets `loss` for every
    game (selfplay.py:86), so only the last game of a batch is differentiated: `reference_coef` builds the
    per-position coefficients of what the reference computes, `intended_coef` those of the sum over all games;
  * torch.optim.AdamW's update (selfplay.py:138: lr 1e-5, betas (0.9, 0.999), eps 1e-8, weight_decay 0.01).
The arithmetic lives in PyTorch CPU kernels + autograd (the reference's own dependency).  Pinned against a run of
the unmodified reference: tests/golden/make_golden_reinforce.py -> tests/golden/reinforce.npz,
checked in tests/test_oracle_golden.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .nets import CONV_IDX, HEAD_IDX, _t

EPS_BN = 1e-5
PROB_EPS = float(torch.finfo(torch.float32).eps)   # Categorical clamps probabilities to [eps, 1-eps] before the log


def param_keys():
    """parameter names in `PolicyNet.parameters()` order (the order AdamW sees them)"""
    keys = []
    for i in CONV_IDX:
        keys += [f"conv.{i}.weight", f"conv.{i}.bias", f"conv.{i + 1}.weight", f"conv.{i + 1}.bias"]
    return keys + [f"conv.{HEAD_IDX}.weight", f"conv.{HEAD_IDX}.bias"]


def train_forward(sd, feats, bn="position"):
    """logits [P,81] of P positions; bn = "position": each position normalised with its own statistics (the reference in
    train mode), "eval": running statistics.  Also returns the per-position statistics a train-mode call would feed into
    the running averages: mean [P,7,128] and UNBIASED variance [P,7,128]."""
    x = _t(feats).float().reshape(-1, 27, 9, 9)
    P = x.shape[0]
    means, uvars = [], []
    for i in CONV_IDX:
        w, b = _t(sd[f"conv.{i}.weight"]), _t(sd[f"conv.{i}.bias"])
        z = F.conv2d(x, w, b, padding=w.shape[-1] // 2)
        g, be = _t(sd[f"conv.{i + 1}.weight"]), _t(sd[f"conv.{i + 1}.bias"])
        means.append(z.detach().mean(dim=(2, 3)))
        uvars.append(z.detach().var(dim=(2, 3), unbiased=True))
        if bn == "position":
            # batch norm over a batch of one == instance norm with the affine parameters
            x = F.instance_norm(z, None, None, g, be, use_input_stats=True, eps=EPS_BN)
        else:
            x = F.batch_norm(z, _t(sd[f"conv.{i + 1}.running_mean"]), _t(sd[f"conv.{i + 1}.running_var"]), g, be,
                             training=False, eps=EPS_BN)
        x = F.relu(x)
    y = F.conv2d(x, _t(sd[f"conv.{HEAD_IDX}.weight"]), None) + _t(sd[f"conv.{HEAD_IDX}.bias"]).unsqueeze(0)
    return y.reshape(P, 81), torch.stack(means, 1), torch.stack(uvars, 1)


def log_prob(logits, moves):
    """Categorical(probs=softmax(logits)).log_prob(move): probs renormalised and clamped to [eps, 1-eps]"""
    p = torch.softmax(logits, dim=1)
    p = p / p.sum(-1, keepdim=True)
    p = p.clamp(min=PROB_EPS, max=1 - PROB_EPS)
    return torch.log(p).gather(1, moves.long().reshape(-1, 1)).reshape(-1)


def reinforce_grads(sd, feats, moves, coef, bn="position"):
    """loss = sum_p coef[p] * (-log_prob_p) and its gradient for every parameter.  Returns (loss, {key: grad}, logits)."""
    keys = param_keys()
    leaf = {k: _t(sd[k]).clone().float().requires_grad_(True) for k in keys}
    full = {k: _t(v) for k, v in sd.items()}
    full.update(leaf)
    logits, _, _ = train_forward(full, feats, bn)
    lp = log_prob(logits, _t(moves))
    loss = (-(lp) * _t(coef).float()).sum()
    loss.backward()
    return float(loss), {k: (leaf[k].grad if leaf[k].grad is not None else torch.zeros_like(leaf[k])) for k in keys}, \
        logits.detach()


def running_stats(sd, means, uvars, seq=None, momentum=0.1):
    """the running_mean / running_var / num_batches_tracked entries after train-mode calls on positions seq[0], seq[1], ...
    (one call per entry, in order)"""
    means, uvars = np.asarray(means, np.float32), np.asarray(uvars, np.float32)
    seq = range(means.shape[0]) if seq is None else seq
    out = {}
    m = np.float32(momentum)
    for li, i in enumerate(CONV_IDX):
        rm = np.asarray(sd[f"conv.{i + 1}.running_mean"], np.float32).copy()
        rv = np.asarray(sd[f"conv.{i + 1}.running_var"], np.float32).copy()
        n = 0
        for p in seq:
            rm = (np.float32(1) - m) * rm + m * means[p, li]
            rv = (np.float32(1) - m) * rv + m * uvars[p, li]
            n += 1
        out[f"conv.{i + 1}.running_mean"], out[f"conv.{i + 1}.running_var"] = rm, rv
        out[f"conv.{i + 1}.num_batches_tracked"] = np.asarray(sd[f"conv.{i + 1}.num_batches_tracked"]).astype(np.int64) + n
    return out


def adamw_step(p, g, m, v, step, lr=1e-5, b1=0.9, b2=0.999, eps=1e-8, wd=0.01):
    """one torch.optim.AdamW update of one tensor (float32 arithmetic, torch's single-tensor formulation);
    `step` counts from 1.  Returns (p, m, v)."""
    p, g, m, v = (torch.as_tensor(np.asarray(a), dtype=torch.float32).clone() for a in (p, g, m, v))
    p.mul_(1 - lr * wd)
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))
    return p, m, v


def replay_positions(lengths, color):
    """(game, ply) of the positions `reinforce` evaluates while replaying a batch (selfplay.py:88-101): the training
    colour's turns of every game, games in order"""
    out = []
    for i, n in enumerate(lengths):
        out += [(i, j) for j in range(1 if color else 0, int(n), 2)]
    return out


def reference_coef(lengths, results, color, bs):
    """per replayed position: what the reference differentiates -- the last game only, reward_last / bs"""
    pos = replay_positions(lengths, color)
    last = len(lengths) - 1
    reward = -results[last] if color else results[last]
    return np.array([reward / bs if g == last else 0.0 for g, _ in pos], np.float32)


def intended_coef(lengths, results, color, bs):
    pos = replay_positions(lengths, color)
    return np.array([(-results[g] if color else results[g]) / bs for g, _ in pos], np.float32)
