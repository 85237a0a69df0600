#!/usr/bin/env python
"""Generates tests/golden/nets_big.npz: the floating-point pin of the conv kernel on a population large enough to state an
error DISTRIBUTION ("re-pin the floating-point bar with data").

Runs the UNMODIFIED reference (imported from /root/reference, CPU, fp32, one thread) -- never part of the product:
    cd /root/repo && python tests/golden/make_golden_nets_big.py

Positions (4,496):
    set 0   400  the survey's own set (SURVEY App. D "precision emulation"): random.seed(0), depth ~ U{0..70}, uniformly
                 random legal moves, evaluated as FRESH go.Game objects
    set 1  2048  every 2nd state of seeded random-legal games (3 % passes), FRESH objects (exact liberties)
    set 2  2048  the same kind of states with the CARRIED liberty cache of the live game object (what the nets
                 see during search / playouts), libs_in = Game._libs before the call
Outputs per position: PolicyNet(policy_17) and PolicyNet(policy_19) logits, stand-in ValueNet value (policy_19 trunk + head
seeded 1234) -- all through the reference's own nn.Modules on nnet.features(game).
"""
import os
import random
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import bokego.go as go        # noqa: E402  (the reference)
import bokego.nnet as nnet    # noqa: E402

from oracle import nets as onets  # noqa: E402

ENC = {go.BLACK: 1, go.WHITE: -1, go.EMPTY: 0}


def state(g):
    return (np.array([ENC[c] for c in g.board], np.int8), -1 if g.ko is None else int(g.ko),
            -2 if g.last_move is None else int(g.last_move), int(g.turn))


def survey_set(n):
    rng = random.Random(0)
    out = []
    while len(out) < n:
        g = go.Game()
        for _ in range(rng.randint(0, 70)):
            legal = g.get_legal_moves()
            if not legal:
                break
            g.play_move(rng.choice(legal))
        out.append(go.Game(g.board, g.ko, g.last_move, g.turn))
    return out


def game_states(seed, want, every=2):
    """(fresh copies, live objects with the carried cache) of states from random-legal games"""
    rng = random.Random(seed)
    fresh, live = [], []
    while len(fresh) < want:
        g = go.Game()
        nnet.features(g)                              # the live object's cache starts at the first call, like in a search
        for t in range(90):
            legal = g.get_legal_moves()
            if not legal or rng.random() < 0.03:
                g.play_move(go.PASS)
            else:
                g.play_move(rng.choice(legal))
            if t % every == 0:
                fresh.append(go.Game(g.board, g.ko, g.last_move, g.turn))
                # snapshot of the live object BEFORE features() refreshes its cache
                snap = go.Game(g.board, g.ko, g.last_move, g.turn)
                snap._libs = bytearray(g._libs)
                live.append(snap)
            nnet.features(g)                          # keep the carried cache moving exactly as mcts.py / selfplay.py do
    return fresh[:want], live[:want]


def main():
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    ck = lambda n: torch.load(os.path.join(REF, "data", "weights", n), map_location="cpu")["model_state_dict"]
    pi17, pi19, vnet = nnet.PolicyNet(), nnet.PolicyNet(), nnet.ValueNet()
    pi17.load_state_dict(ck("policy_17.pt")); pi19.load_state_dict(ck("policy_19.pt"))
    vnet.load_policy_dict(pi19.state_dict())
    sd = vnet.state_dict(); sd.update(onets.standin_value_head(1234)); vnet.load_state_dict(sd)
    for m in (pi17, pi19, vnet):
        m.eval()
    s0 = survey_set(400)
    s1, s2 = game_states(7, 2048)
    games = [(g, 0) for g in s0] + [(g, 1) for g in s1] + [(g, 2) for g in s2]
    rows = {k: [] for k in ("board", "ko", "last", "turn", "set", "libs_in")}
    feats = []
    for g, tag in games:
        b, ko, last, turn = state(g)
        rows["board"].append(b); rows["ko"].append(ko); rows["last"].append(last); rows["turn"].append(turn)
        rows["set"].append(tag)
        rows["libs_in"].append(np.zeros(81, np.uint8) if g._libs is None else np.frombuffer(bytes(g._libs), np.uint8).copy())
        feats.append(nnet.features(g))
    x = torch.stack(feats)
    with torch.no_grad():
        l17 = torch.cat([pi17(x[i:i + 512]) for i in range(0, len(x), 512)])
        l19 = torch.cat([pi19(x[i:i + 512]) for i in range(0, len(x), 512)])
        val = torch.cat([vnet(x[i:i + 512]).reshape(-1) for i in range(0, len(x), 512)])
    np.savez_compressed(os.path.join(HERE, "nets_big.npz"), board=np.stack(rows["board"]), ko=np.array(rows["ko"], np.int16),
                        last=np.array(rows["last"], np.int16), turn=np.array(rows["turn"], np.int16),
                        set=np.array(rows["set"], np.uint8), libs_in=np.stack(rows["libs_in"]),
                        logits17=l17.numpy(), logits19=l19.numpy(), value=val.numpy(), value_head_seed=1234)
    n_stale = sum(1 for (g, t), f in zip(games, feats) if t == 2 and not torch.equal(
        f, nnet.features(go.Game(g.board, g.ko, g.last_move, g.turn))))
    print(f"positions {len(games)} (sets 400 / {len(s1)} / {len(s2)}); carried-cache positions whose planes differ from fresh: {n_stale}")
    print(f"nets_big.npz: {os.path.getsize(os.path.join(HERE, 'nets_big.npz')) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
