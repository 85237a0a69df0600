#!/usr/bin/env python
"""Golden fixture for the REINFORCE step made by RUNNING THE UNMODIFIED REFERENCE.

Runs only in the build container (needs /root/reference):

    python tests/golden/make_golden_reinforce.py

It calls the reference's own `reinforce` (bin/selfplay.py:59-122) for one iteration per case with
`pi` = policy_17 in train() mode (as bin/selfplay.py:148-150 sets it), `pi_opp` = policy_19 in eval() mode and
torch.optim.AdamW(lr=1e-5) (selfplay.py:138), and records, per case,

  * the games and results its self-play produced, and the stream of feature planes that went through `pi`
    in call order (forward pre-hook) -- every call is a training-mode BatchNorm batch of ONE position
    (nnet.py:265-275), so this stream is what the running statistics are filtered over;
  * which positions / moves / rewards enter the loss.  The reference resets `loss` per game
    (selfplay.py:86) so only the LAST game of the batch reaches `loss.backward()`; the fixture keeps
    that behaviour (it is what the reference computes) and the tests also exercise the intended sum;
  * `p.grad` of every parameter after the iteration, the parameters after `optimizer.step()`, and the
    BatchNorm running statistics -- small tensors in full, conv weights as a strided sample plus
    float64 sum / L2 norm over the whole tensor.

Shims: `go.gnu_score` needs the gnugo binary, which is absent; it is replaced by the
sign of Game.score(), the stand-in the rest of this repo uses.  Nothing else is patched.
"""
import copy
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "bin"))
HERE = os.path.dirname(os.path.abspath(__file__))

import bokego.go as go            # noqa: E402  (the reference)
import bokego.nnet as nnet        # noqa: E402
import selfplay                   # noqa: E402  (reference bin/selfplay.py)

torch.set_num_threads(1)
STRIDE = 32                       # sampling stride for the conv weights


def load(n):
    net = nnet.PolicyNet()
    ck = torch.load(os.path.join(REF, "data", "weights", f"policy_{n}.pt"), map_location="cpu")
    net.load_state_dict(ck["model_state_dict"])
    return net


def compact(name, t, out, prefix):
    a = t.detach().cpu().numpy().astype(np.float32).ravel()
    out[f"{prefix}/{name}/sum"] = np.float64(a.astype(np.float64).sum())
    out[f"{prefix}/{name}/l2"] = np.float64(np.sqrt((a.astype(np.float64) ** 2).sum()))
    out[f"{prefix}/{name}/absmax"] = np.float64(np.abs(a).max())
    out[f"{prefix}/{name}"] = a if a.size <= 4096 else a[::STRIDE].copy()


def run_case(tag, color, bs, seed, out):
    pi, opp = load(17), load(19)
    pi.train()
    opp.eval()
    opt = torch.optim.AdamW(pi.parameters(), lr=1e-5)
    calls, marks = [], {}
    pi.register_forward_pre_hook(lambda m, inp: calls.append(inp[0][0].to(torch.uint8).numpy().copy()))
    go.gnu_score = lambda g: 1 if g.score() > 0 else -1
    rec = {}
    orig = selfplay.self_play

    def recorder(p1, p2, n, device=None):
        games, results = orig(p1, p2, n, device=device)
        rec["games"], rec["results"] = copy.deepcopy(games), list(results)
        marks["replay_from"] = len(calls)
        return games, results

    selfplay.self_play = recorder
    torch.manual_seed(seed)
    stats = []
    try:
        selfplay.reinforce(pi, opp, opt, color, n_itrs=1, bs=bs, device=torch.device("cpu"), stats=stats, id=0)
    finally:
        selfplay.self_play = orig
    L = max(len(g) for g in rec["games"])
    moves = np.full((bs, L), -9, np.int16)
    for i, g in enumerate(rec["games"]):
        moves[i, :len(g)] = g
    out[f"{tag}/color"] = np.array(0 if color == "black" else 1)
    out[f"{tag}/bs"] = np.array(bs)
    out[f"{tag}/moves"] = moves
    out[f"{tag}/lengths"] = np.array([len(g) for g in rec["games"]], np.int32)
    out[f"{tag}/results"] = np.array(rec["results"], np.int32)
    out[f"{tag}/wins"] = np.array(stats, np.int32)
    out[f"{tag}/calls"] = np.stack(calls).reshape(len(calls), 27, 81)
    out[f"{tag}/replay_from"] = np.array(marks["replay_from"])
    for k, p in pi.named_parameters():
        compact(k, p.grad, out, f"{tag}/grad")
    for k, v in pi.state_dict().items():
        if k.endswith("num_batches_tracked"):
            out[f"{tag}/post/{k}"] = v.numpy()
        else:
            compact(k, v, out, f"{tag}/post")
    print(tag, "games", [len(g) for g in rec["games"]], "results", rec["results"], "calls", len(calls),
          "replay_from", marks["replay_from"], "wins", stats)


def main():
    out = {"stride": np.array(STRIDE)}
    run_case("black3", "black", 3, 11, out)
    run_case("white2", "white", 2, 12, out)
    np.savez_compressed(os.path.join(HERE, "reinforce.npz"), **out)
    print("wrote reinforce.npz", os.path.getsize(os.path.join(HERE, "reinforce.npz")))


if __name__ == "__main__":
    main()
