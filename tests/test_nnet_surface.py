"""bokego_b200.nnet: module surface and state-dict compatibility with the reference (CPU only; no compute)."""
import numpy as np
import pytest
import torch

from bokego_b200 import _lib, nnet


def test_names_resolve():
    for name in ("PolicyNet", "ValueNet", "PolicyNet_v2", "Conv2dUntiedBias", "SOFT", "features", "policy_dist", "value",
                 "policy_sample", "features_batch", "policy_value_batch", "playout_step", "score_batch", "prefill_caches"):
        assert hasattr(nnet, name), name
    assert isinstance(nnet.SOFT, torch.nn.Softmax) and nnet.SOFT.dim == 1


def test_state_dict_keys_match_reference_checkpoints(sd17, sd_value):
    pi, v = nnet.PolicyNet(), nnet.ValueNet()
    assert set(pi.state_dict().keys()) == set(sd17.keys())
    for k, t in pi.state_dict().items():
        assert tuple(t.shape) == tuple(sd17[k].shape), k
    assert set(v.state_dict().keys()) == set(sd_value.keys())
    pi.load_state_dict({k: torch.from_numpy(np.asarray(a)) for k, a in sd17.items()})      # strict load works
    v.load_policy_dict(pi.state_dict())
    assert torch.equal(v.conv[0].weight, pi.conv[0].weight) and torch.equal(v.conv[21].bias, pi.conv[21].bias)
    assert sum(p.numel() for p in pi.parameters()) == 974033
    u = nnet.Conv2dUntiedBias(9, 9, 128, 1, 1)
    assert tuple(u.weight.shape) == (1, 128, 1, 1) and tuple(u.bias.shape) == (1, 9, 9)
    with pytest.raises(ValueError):
        nnet.Conv2dUntiedBias(9, 9, 3, 2, 1, groups=2)
    assert len(nnet.PolicyNet_v2().state_dict()) == 14


def test_no_cpu_path():
    """the product path must fail loudly without a CUDA device / on CPU tensors"""
    pi = nnet.PolicyNet().eval()
    with pytest.raises(_lib.BokegoB200Error):
        pi(torch.zeros(1, 27, 9, 9))
    with pytest.raises(_lib.BokegoB200Error):
        _lib.require_device(torch.device("cpu"))


def test_c_abi_exports_every_declared_symbol():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "bokego_b200.h")).read()
    names = set(re.findall(r"\b(bk_[a-z0-9_]+)\s*\(", hdr))
    assert {"bk_encode", "bk_forward", "bk_playout_step", "bk_score", "bk_weights_pack", "bk_repack_f32"} <= names
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n), n
    assert L.bk_version() >= 100 and L.bk_strerror(-1) == b"bad argument"
