"""GPU tests of the drop-in surface: bokego_b200.nnet / bokego_b200.go driven the way the reference's
mcts.py / selfplay.py drive bokego.nnet / bokego.go (one position at a time), plus the batched cache prefill."""
import numpy as np
import pytest
import torch

from bokego_b200 import _lib, go, nnet

pytestmark = pytest.mark.gpu
CH = {1: go.BLACK, -1: go.WHITE, 0: go.EMPTY}
DEV = torch.device("cuda", 0)


def _game(p, i, cls=go.Game):
    g = cls("".join(CH[int(v)] for v in p["board"][i]), None if p["ko"][i] < 0 else int(p["ko"][i]),
            None if p["last"][i] == -2 else int(p["last"][i]), int(p["turn"][i]))
    if not p["fresh"][i]:
        g._libs = bytearray(p["libs_in"][i].tobytes())
    return g


def _nets(sd17, sd_value):
    pi, v = nnet.PolicyNet(), nnet.ValueNet()
    pi.load_state_dict({k: torch.from_numpy(np.asarray(a)) for k, a in sd17.items()})
    v.load_state_dict({k: torch.from_numpy(np.asarray(a)) for k, a in sd_value.items()})
    return pi.eval().to(DEV), v.eval().to(DEV)


def test_features_single_position(positions):
    p = positions
    for i in list(range(0, len(p["board"]), 97)) + list(range(len(p["board"]) - 6, len(p["board"]))):
        g = _game(p, i)
        f = nnet.features(g)
        assert f.dtype == torch.float32 and tuple(f.shape) == (27, 9, 9) and f.device.type == "cpu"
        assert np.array_equal(f.numpy().reshape(27, 81), p["feats"][i].astype(np.float32)), i
        assert bytes(g._libs) == p["libs_out"][i].tobytes()


def test_nets_and_wrappers(positions, nets_golden, sd17, sd_value):
    pi, v = _nets(sd17, sd_value)
    src = nets_golden["src"][:64]
    x = torch.from_numpy(positions["feats"][src]).float().reshape(-1, 27, 9, 9).to(DEV)
    logits, val = pi(x), v(x)
    assert tuple(logits.shape) == (64, 81) and tuple(val.shape) == (64, 1)
    assert float((logits.cpu() - torch.from_numpy(nets_golden["logits17"][:64])).abs().max()) < 5e-2
    assert float((val.cpu().reshape(-1) - torch.from_numpy(nets_golden["value"][:64])).abs().max()) < 1e-3
    g = _game(positions, int(src[3]))
    d = nnet.policy_dist(pi, g, device=DEV)
    assert abs(float(d.probs.sum()) - 1) < 1e-5 and int(d.probs.argmax()) == int(nets_golden["logits17"][3].argmax())
    assert abs(nnet.value(v, g, device=DEV) - float(nets_golden["value"][3])) < 1e-3
    mv = nnet.policy_sample(pi, g, device=DEV)
    assert mv.dim() == 0 and mv.dtype == torch.int64 and 0 <= int(mv) < 81
    # reload weights -> the packed blob is refreshed
    pi.load_state_dict({k: torch.from_numpy(np.asarray(a)) * (0.5 if k == "conv.21.bias" else 1.0) for k, a in sd17.items()})
    l2 = pi(x[:1])
    assert float((l2 - logits[:1]).abs().max()) > 1e-3
    # train() mode: one position per call, normalised with its own statistics, running averages filtered like torch does
    # (how policy_dist calls a net that bin/selfplay.py:148-150 has put in train mode); batches and the ValueNet raise
    from oracle import train as ot
    pi.load_state_dict({k: torch.from_numpy(np.asarray(a)) for k, a in sd17.items()})
    pi.train()
    with pytest.raises(_lib.BokegoB200Error):
        pi(x[:2])
    v.train()
    with pytest.raises(_lib.BokegoB200Error):
        v(x[:1])
    nbt = int(pi.state_dict()["conv.1.num_batches_tracked"])
    lt = pi(x[5:6])
    want, means, uvars = ot.train_forward(sd17, x[5:6].cpu())
    assert float((lt.cpu() - want).abs().max()) < 2e-3
    assert float((lt.cpu() - logits[5:6].cpu()).abs().max()) > 1e-2          # not the eval-mode function
    rs = ot.running_stats(sd17, means.numpy(), uvars.numpy())
    sd_after = pi.state_dict()
    for k, w in rs.items():
        if k.endswith("tracked"):
            assert int(sd_after[k]) == nbt + 1
        else:
            assert float((sd_after[k].cpu() - torch.from_numpy(w)).abs().max()) < 1e-3 * (1 + float(np.abs(w).max())), k
    d = nnet.policy_dist(pi, g, device=DEV)                                   # the reference's call, net in train mode
    assert abs(float(d.probs.sum()) - 1) < 1e-5 and int(sd_after["conv.1.num_batches_tracked"]) == nbt + 2


def test_prefill_caches(positions, nets_golden, sd17, sd_value):
    pi, v = _nets(sd17, sd_value)

    class Tree:                       # the three class-level caches of the reference's MCTS (mcts.py:42-44)
        _val_cache, _dist_cache, _fts_cache = dict(), dict(), dict()

    class Node(go.Game):
        def __hash__(self):
            return super().__hash__()

        def __eq__(self, o):
            return self.board == o.board and self.ko == o.ko and self.last_move == o.last_move

    src = nets_golden["src"][:40]
    carried = [i for i in range(len(positions["board"])) if not positions["fresh"][i]][:23]
    nodes = [_game(positions, int(i), Node) for i in list(src) + carried]
    n = nnet.prefill_caches(Tree, nodes, pi, v, device=DEV)
    assert n == len(nodes) and nnet.prefill_caches(Tree, nodes, pi, v, device=DEV) == 0
    for k, i in enumerate(list(src) + carried):
        nd = nodes[k]
        assert np.array_equal(Tree._fts_cache[nd].numpy().reshape(27, 81), positions["feats"][i].astype(np.float32))
        assert bytes(nd._libs) == positions["libs_out"][i].tobytes()
        assert abs(float(Tree._dist_cache[nd].probs.sum()) - 1) < 1e-5 and isinstance(Tree._val_cache[nd], float)
    for k in range(8):
        assert int(Tree._dist_cache[nodes[k]].probs.argmax()) == int(nets_golden["logits17"][k].argmax())
        assert abs(Tree._val_cache[nodes[k]] - float(nets_golden["value"][k])) < 1e-3
