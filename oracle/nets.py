"""fp32 CPU restatement of the reference nets' forward pass.  TEST INFRASTRUCTURE ONLY.

Functional (state-dict in, tensors out) restatement of
  PolicyNet.forward  /root/reference/bokego/nnet.py:19-57
  ValueNet.forward   /root/reference/bokego/nnet.py:59-113
  Conv2dUntiedBias   /root/reference/bokego/nnet.py:138-180
  SOFT               /root/reference/bokego/nnet.py:16
in eval mode (BatchNorm uses running statistics, eps = 1e-5).  The arithmetic itself lives in
PyTorch's CPU kernels (torch is the reference's own, unpinned, dependency -- setup.py:16); the pin
is "torch 2.11 CPU fp32 in this image".  Checked against the real nn.Modules in
tests/golden/make_golden.py (logits agree to ~1e-5, the batch-vs-single self-noise of torch itself).
"""
import numpy as np
import torch
import torch.nn.functional as F

CONV_IDX = (0, 3, 6, 9, 12, 15, 18)   # positions of the Conv2d layers inside `conv` (nnet.py:31-52)
HEAD_IDX = 21                          # Conv2dUntiedBias


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.asarray(x))


def trunk(sd, x):
    """conv.0 .. conv.20: 7 x (conv -> BN(eval) -> ReLU); x float32 [B,27,9,9] -> [B,128,9,9]."""
    for i in CONV_IDX:
        w, b = _t(sd[f"conv.{i}.weight"]), _t(sd[f"conv.{i}.bias"])
        x = F.conv2d(x, w, b, padding=w.shape[-1] // 2)
        j = i + 1
        x = F.batch_norm(x, _t(sd[f"conv.{j}.running_mean"]), _t(sd[f"conv.{j}.running_var"]),
                         _t(sd[f"conv.{j}.weight"]), _t(sd[f"conv.{j}.bias"]), training=False, eps=1e-5)
        x = F.relu(x)
    return x


def head81(sd, x):
    """conv.21: 1x1 conv 128->1 without bias, plus the per-square bias (1,9,9) -> [B,81]."""
    y = F.conv2d(x, _t(sd[f"conv.{HEAD_IDX}.weight"]), None) + _t(sd[f"conv.{HEAD_IDX}.bias"]).unsqueeze(0)
    return y.reshape(-1, 81)


@torch.no_grad()
def policy_logits(sd, feats):
    return head81(sd, trunk(sd, _t(feats).float()))


@torch.no_grad()
def policy_probs(sd, feats):
    return torch.softmax(policy_logits(sd, feats), dim=1)


@torch.no_grad()
def value(sd, feats):
    """ValueNet.forward -> [B]."""
    h = head81(sd, trunk(sd, _t(feats).float())).reshape(-1, 1, 9, 9)
    h = F.batch_norm(h, _t(sd["bn.running_mean"]), _t(sd["bn.running_var"]), _t(sd["bn.weight"]),
                     _t(sd["bn.bias"]), training=False, eps=1e-5)
    h = F.relu(h).reshape(-1, 81)
    h = F.linear(h, _t(sd["lin1.weight"]), _t(sd["lin1.bias"]))
    h = F.batch_norm(h, _t(sd["lin_bn.running_mean"]), _t(sd["lin_bn.running_var"]), _t(sd["lin_bn.weight"]),
                     _t(sd["lin_bn.bias"]), training=False, eps=1e-5)
    h = F.relu(h)
    return torch.tanh(F.linear(h, _t(sd["lin2.weight"]), _t(sd["lin2.bias"]))).reshape(-1)


def folded(sd):
    """BatchNorm(eval) folded into the seven convs in float64: w' = w * g / sqrt(var + eps), b' = (b - mean) * g / sqrt(var + eps) + beta"""
    ws, bs = [], []
    for i in CONV_IDX:
        w, b = _t(sd[f"conv.{i}.weight"]).double(), _t(sd[f"conv.{i}.bias"]).double()
        s = _t(sd[f"conv.{i + 1}.weight"]).double() / torch.sqrt(_t(sd[f"conv.{i + 1}.running_var"]).double() + 1e-5)
        ws.append(w * s[:, None, None, None])
        bs.append((b - _t(sd[f"conv.{i + 1}.running_mean"]).double()) * s + _t(sd[f"conv.{i + 1}.bias"]).double())
    return ws, bs


@torch.no_grad()
def policy_logits_f16_operands(sd, feats):
    """The arithmetic the north-star prescribes for the conv kernel, emulated on the CPU: BatchNorm folded, conv weights and
    the activations handed from layer to layer rounded to fp16 (round-to-nearest-even), every sum in fp32, bias / ReLU / the
    1x1 head in fp32.  NOT the reference's result: it is the yardstick that separates "error of 16-bit operands" (what this
    function shows against policy_logits) from "error of the kernel" (what the kernel shows against this function)."""
    h = lambda t: t.float().half().float()
    ws, bs = folded(sd)
    a = _t(feats).float()
    for l in range(7):
        y = F.conv2d(a, h(ws[l]), None, padding=ws[l].shape[-1] // 2) + bs[l].float()[None, :, None, None]
        y = F.relu(y)
        a = h(y) if l < 6 else y
    return head81(sd, a)


def planes_to_float(feats_u8):
    """uint8 [B,27,81] plane values -> float32 [B,27,9,9] exactly as nnet.features returns them."""
    return torch.from_numpy(np.asarray(feats_u8)).float().reshape(-1, 27, 9, 9)


def standin_value_head(seed=1234):
    """Seeded head parameters for the stand-in ValueNet (the reference's value_1.pt is not shipped,
    SURVEY F3).  Non-trivial BatchNorm statistics so that folding errors would show."""
    g = torch.Generator().manual_seed(seed)
    u = lambda *s, lo=-1.0, hi=1.0: (torch.rand(*s, generator=g) * (hi - lo) + lo)
    return {
        "lin1.weight": u(64, 81) / 9.0, "lin1.bias": u(64) / 9.0,
        "lin2.weight": u(1, 64) / 8.0, "lin2.bias": u(1) / 8.0,
        "bn.weight": u(1, lo=0.5, hi=1.5), "bn.bias": u(1, lo=-0.2, hi=0.2),
        "bn.running_mean": u(1, lo=-3.0, hi=1.0), "bn.running_var": u(1, lo=4.0, hi=30.0),
        "bn.num_batches_tracked": torch.tensor(1),
        "lin_bn.weight": u(64, lo=0.5, hi=1.5), "lin_bn.bias": u(64, lo=-0.3, hi=0.3),
        "lin_bn.running_mean": u(64, lo=-0.5, hi=0.5), "lin_bn.running_var": u(64, lo=0.5, hi=2.0),
        "lin_bn.num_batches_tracked": torch.tensor(1),
    }
