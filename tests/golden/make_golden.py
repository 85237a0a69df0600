#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Runs only in the build container (needs /root/reference; the GPU box has no such path):

    python tests/golden/make_golden.py

The reference (meiji163/bokego) is imported from /root/reference via sys.path -- nothing is copied.
Everything the tests compare against comes out of reference code executed here:

  weights_policy_{17,19}.npz  state dicts of data/weights/policy_{17,19}.pt (fp32, key names kept)
  positions.npz               states -> nnet.features / legal / _libs, in both liberty-cache modes:
                              SGF replays (data/bokevgnugo), seeded random-legal games with passes,
                              and the known-answer quirk positions of SURVEY F5/F6/F7
  rules.npz                   per state: play_move outcome for all 81 squares, Game.is_legal,
                              go.possible_eye, Game.score
  nets.npz                    PolicyNet(policy_17/19) logits and stand-in ValueNet values (fp32 CPU)
  playouts.npz                move-by-move traces of Go_MCTS.find_random_child (mcts flavour) and
                              bin/selfplay.legal_sample (self-play flavour) with the random draws
                              recorded in the form ATen consumes them (81 Exp(1) variates per draw)

Shims applied to the reference while tracing, none of which changes a result the
reference itself produces:
  * Categorical.sample is replaced by argmax(probs / q), q = torch.empty(81).exponential_(1): this
    IS torch.multinomial's single-sample algorithm; the script asserts that it reproduces the
    unshimmed sampler draw for draw under the same seed.
  * get_move: when no probability mass is left the reference raises inside torch.multinomial
    (SURVEY F7); the trace records PASS, the documented intent of mcts.py:349-350.
"""
import hashlib
import os
import random
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "bin"))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import bokego.go as go            # noqa: E402  (the reference)
import bokego.nnet as nnet        # noqa: E402
import bokego.mcts as mcts        # noqa: E402
import selfplay                   # noqa: E402  (reference bin/selfplay.py)
from torch.distributions.categorical import Categorical  # noqa: E402

from oracle import nets as onets  # noqa: E402

torch.set_num_threads(1)
ENC = {go.BLACK: 1, go.WHITE: -1, go.EMPTY: 0}


def board_arr(b):
    return np.array([ENC[c] for c in b], dtype=np.int8)


def enc_ko(k):
    return -1 if k is None else int(k)


def enc_last(l):
    return -2 if l is None else int(l)


class Rec:
    """accumulates (state -> features) records"""

    def __init__(self):
        self.rows = {k: [] for k in ("board", "ko", "last", "turn", "fresh", "libs_in", "feats", "legal", "libs_out", "tag")}

    def add(self, g, tag):
        """call nnet.features(g) on the object as it is (incremental cache if it has one)"""
        fresh = g._libs is None
        libs_in = np.zeros(81, np.uint8) if fresh else np.frombuffer(bytes(g._libs), np.uint8).copy()
        st = (board_arr(g.board), enc_ko(g.ko), enc_last(g.last_move), g.turn)
        f = nnet.features(g).numpy()
        assert f.shape == (27, 9, 9) and np.all(f == np.round(f)) and f.min() >= 0 and f.max() <= 7
        legal = np.zeros(81, np.uint8)
        legal[g.get_legal_moves()] = 1
        r = self.rows
        r["board"].append(st[0]); r["ko"].append(st[1]); r["last"].append(st[2]); r["turn"].append(st[3])
        r["fresh"].append(int(fresh)); r["libs_in"].append(libs_in)
        r["feats"].append(f.reshape(27, 81).astype(np.uint8)); r["legal"].append(legal)
        r["libs_out"].append(np.frombuffer(bytes(g._libs), np.uint8).copy()); r["tag"].append(tag)
        return f

    def save(self, path):
        r = self.rows
        np.savez_compressed(path, board=np.stack(r["board"]), ko=np.array(r["ko"], np.int16),
                            last=np.array(r["last"], np.int16), turn=np.array(r["turn"], np.int16),
                            fresh=np.array(r["fresh"], np.uint8), libs_in=np.stack(r["libs_in"]),
                            feats=np.stack(r["feats"]), legal=np.stack(r["legal"]),
                            libs_out=np.stack(r["libs_out"]), tag=np.array(r["tag"], np.int16))


def random_game_states(seed, n_games, n_moves, pass_prob):
    """seeded uniformly-random legal play (SURVEY 8d generator); yields live Game objects"""
    rng = random.Random(seed)
    for _ in range(n_games):
        g = go.Game()
        yield g
        for _ in range(n_moves):
            legal = g.get_legal_moves()
            if not legal or rng.random() < pass_prob:
                g.play_move(go.PASS)
            else:
                g.play_move(rng.choice(legal))
            yield g


def make_positions():
    rec = Rec()
    # (1) SGF replays, incremental cache -- also re-derive the SHA-256 prefixes quoted in SURVEY App. D
    sha = []
    for i in range(1, 11):
        g = go.Game(sgf=os.path.join(REF, "data", "bokevgnugo", f"boke_gnugo_{i}.sgf"))
        h = hashlib.sha256()
        for _ in range(len(g.moves)):
            g.play_move()
            h.update(rec.add(g, tag=i).astype(np.float32).tobytes())
        sha.append(h.hexdigest()[:12])
    print("sgf feature sha256 prefixes:", " ".join(sha))
    # (2) random-legal games, with passes; every state in incremental mode and as a fresh object
    n_inc = n_diff = 0
    for g in random_game_states(seed=1, n_games=24, n_moves=90, pass_prob=0.03):
        f_inc = rec.add(g, tag=100)
        f_fr = rec.add(go.Game(g.board, g.ko, g.last_move, g.turn), tag=101)
        n_inc += 1
        n_diff += int(not np.array_equal(f_inc, f_fr))
    print(f"random games: {n_inc} states, {n_diff} differ between incremental and fresh liberties")
    # (3) known-answer quirk positions
    rows = [".XX......", "XOOX.....", "XO.X.....", ".XX......"] + ["........."] * 5   # F5: double count
    rec.add(go.Game("".join(rows), turn=0), tag=200)
    rec.add(go.Game("XXXX....O" + "XXXX....." * 8, turn=1), tag=201)                 # F6 score board
    rec.add(go.Game(".X......O" + ".X......." * 8, turn=0), tag=202)
    allx = "." + "X" * 79 + "."                                                        # F7 board
    rec.add(go.Game(allx, turn=0), tag=203)
    rec.add(go.Game(allx, turn=1), tag=204)
    rec.add(go.Game("O" * 40 + "." + "X" * 40, turn=0), tag=205)
    rec.save(os.path.join(HERE, "positions.npz"))
    print("positions:", len(rec.rows["board"]))
    return rec


def make_rules(rec):
    """play_move for every square, is_legal, possible_eye, score on a spread of states"""
    r = rec.rows
    n = len(r["board"])
    idx = [i for i in range(n) if r["ko"][i] >= 0][:60]
    rng = random.Random(7)
    idx += rng.sample(range(n), 200) + list(range(n - 6, n))
    idx = sorted(set(idx))
    chars = {1: go.BLACK, -1: go.WHITE, 0: go.EMPTY}
    out = {k: [] for k in ("src", "status", "nboard", "nko", "islegal", "eye", "score", "pass_state")}
    code = {"ko": 1, "not_empty": 2, "suicide": 3}
    for i in idx:
        bstr = "".join(chars[int(v)] for v in r["board"][i])
        ko = None if r["ko"][i] < 0 else int(r["ko"][i])
        last = None if r["last"][i] == -2 else int(r["last"][i])
        turn = int(r["turn"][i])
        status = np.zeros(81, np.int8); nko = np.full(81, -1, np.int16)
        nboard = np.zeros((81, 81), np.int8); isl = np.zeros(81, np.uint8); eye = np.zeros(81, np.int8)
        for s in range(81):
            g = go.Game(bstr, ko, last, turn)
            isl[s] = int(g.is_legal(s))
            e = go.possible_eye(bstr, s)
            eye[s] = 0 if e is None else ENC[e]
            try:
                g.play_move(s)
                nboard[s] = board_arr(g.board); nko[s] = enc_ko(g.ko)
                assert g.turn == turn + 1 and g.last_move == s
            except go.IllegalMove as ex:
                status[s] = code[ex.rule_type]; nboard[s] = r["board"][i]
        g = go.Game(bstr, ko, last, turn)
        g.play_move(go.PASS)
        out["pass_state"].append((enc_ko(g.ko), enc_last(g.last_move), g.turn))
        out["src"].append(i); out["status"].append(status); out["nboard"].append(nboard); out["nko"].append(nko)
        out["islegal"].append(isl); out["eye"].append(eye)
        out["score"].append(go.Game(bstr, ko, last, turn).score())
    np.savez_compressed(os.path.join(HERE, "rules.npz"), src=np.array(out["src"], np.int32),
                        status=np.stack(out["status"]), nboard=np.stack(out["nboard"]), nko=np.stack(out["nko"]),
                        islegal=np.stack(out["islegal"]), eye=np.stack(out["eye"]),
                        score=np.array(out["score"], np.float64), pass_state=np.array(out["pass_state"], np.int16))
    print("rules: states", len(idx), "ko states", sum(1 for i in idx if r["ko"][i] >= 0))


def load_policy(name):
    ck = torch.load(os.path.join(REF, "data", "weights", name), map_location="cpu")
    pi = nnet.PolicyNet()
    pi.load_state_dict(ck["model_state_dict"])
    pi.eval()
    return pi


def save_weights(pi, name):
    np.savez_compressed(os.path.join(HERE, name), **{k: v.numpy() for k, v in pi.state_dict().items()})


def make_nets(rec, pi17, pi19):
    r = rec.rows
    fresh_idx = [i for i in range(len(r["board"])) if r["fresh"][i] and r["tag"][i] == 101][::7][:256]
    x = torch.from_numpy(np.stack([r["feats"][i] for i in fresh_idx])).float().reshape(-1, 27, 9, 9)
    vnet = nnet.ValueNet()
    vnet.load_policy_dict(pi19.state_dict())
    sd = vnet.state_dict(); sd.update(onets.standin_value_head(1234)); vnet.load_state_dict(sd); vnet.eval()
    with torch.no_grad():
        l17, l19, val = pi17(x), pi19(x), vnet(x).reshape(-1)
        # oracle restatement vs the real modules
        o17 = onets.policy_logits({k: v for k, v in pi17.state_dict().items()}, x)
        ov = onets.value({k: v for k, v in vnet.state_dict().items()}, x)
    print("oracle nets vs reference modules: logits", float((o17 - l17).abs().max()), "value", float((ov - val).abs().max()))
    assert float((o17 - l17).abs().max()) < 1e-4 and float((ov - val).abs().max()) < 1e-5
    np.savez_compressed(os.path.join(HERE, "nets.npz"), src=np.array(fresh_idx, np.int32),
                        logits17=l17.numpy(), logits19=l19.numpy(), value=val.numpy(), value_head_seed=1234)
    print("nets:", len(fresh_idx), "positions; value range", float(val.min()), float(val.max()))
    return vnet


# ---- playout traces ---------------------------------------------------------------------------------
class NoMass(Exception):
    pass


class DrawTap:
    """replaces Categorical.sample by the exponential race with recorded draws"""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)
        self.draws, self.probs0 = [], None

    def sample(self, dist, sample_shape=torch.Size()):
        p = dist.probs
        if self.probs0 is None:
            self.probs0 = p.detach().clone().numpy()
        if float(p.sum()) <= 0:
            raise NoMass()
        q = torch.empty(81).exponential_(1, generator=self.gen)
        self.draws.append(q.numpy().copy())
        return torch.argmax(p / q)

    def take(self):
        d, p = self.draws, self.probs0
        self.draws, self.probs0 = [], None
        return np.stack(d), p


def check_sampler_shim():
    """the shim equals the unshimmed sampler under the same seed (incl. zeroed entries)"""
    ok = 0
    for s in range(300):
        torch.manual_seed(s)
        p = torch.softmax(torch.randn(81) * 3, 0)
        p[torch.randperm(81)[: s % 40]] = 0
        d = Categorical(p)
        torch.manual_seed(1000 + s); a = int(d.sample())
        torch.manual_seed(1000 + s); q = torch.empty(81).exponential_(1); b = int(torch.argmax(d.probs / q))
        ok += int(a == b)
    print("sampler shim == torch sampler:", ok, "/ 300")
    assert ok == 300


class FakeTree:
    def __init__(self, pi):
        self.policy_net, self.value_net, self.device = pi, None, torch.device("cpu")


def trace_mcts(pi, n_playouts, seed0):
    T = {k: [] for k in ("board", "ko", "last", "turn", "libs_in", "fresh", "probs", "q", "nq", "move", "game")}
    finals = []
    orig = Categorical.sample
    for gi in range(n_playouts):
        tap = DrawTap(seed0 + gi)
        Categorical.sample = lambda self, sample_shape=torch.Size(), _t=tap: _t.sample(self, sample_shape)
        for c in (mcts.MCTS._dist_cache, mcts.MCTS._fts_cache, mcts.MCTS._val_cache):
            c.clear()
        node = mcts.Go_MCTS()
        node.tree = FakeTree(pi)
        while not node._terminal:
            fresh = node._libs is None
            st = dict(board=board_arr(node.board), ko=enc_ko(node.ko), last=enc_last(node.last_move), turn=node.turn,
                      libs_in=np.zeros(81, np.uint8) if fresh else np.frombuffer(bytes(node._libs), np.uint8).copy(),
                      fresh=int(fresh))
            try:
                mv = node.get_move()
            except NoMass:
                mv = go.PASS                                   # F7 shim
            q, p0 = tap.take()
            child = node.make_move(mv)
            for k, v in st.items():
                T[k].append(v)
            qq = np.zeros((82, 81), np.float32); qq[: len(q)] = q
            T["probs"].append(p0); T["q"].append(qq); T["nq"].append(len(q)); T["move"].append(mv); T["game"].append(gi)
            node = child
        finals.append((board_arr(node.board), node.turn, enc_last(node.last_move), 1 if node.score() > 0 else -1,
                       node.score()))
    Categorical.sample = orig
    return T, finals


def trace_selfplay(pi1, pi2, n_games, seed0):
    T = {k: [] for k in ("board", "ko", "last", "turn", "libs_in", "fresh", "probs", "q", "move", "game", "sampled_legal")}
    finals = []
    orig = Categorical.sample
    cpu = torch.device("cpu")
    for gi in range(n_games):
        tap = DrawTap(seed0 + gi)
        Categorical.sample = lambda self, sample_shape=torch.Size(), _t=tap: _t.sample(self, sample_shape)
        g = go.Game(moves=[])
        stop = False
        while not stop:
            if g.turn > selfplay.MAX_TURNS:
                break
            for pi in (pi1, pi2):
                fresh = g._libs is None
                st = dict(board=board_arr(g.board), ko=enc_ko(g.ko), last=enc_last(g.last_move), turn=g.turn,
                          libs_in=np.zeros(81, np.uint8) if fresh else np.frombuffer(bytes(g._libs), np.uint8).copy(),
                          fresh=int(fresh))
                mv = selfplay.legal_sample(pi, g, cpu)
                q, p0 = tap.take()
                assert len(q) == 1
                sampled = int(np.argmax(p0 / q[0]))
                for k, v in st.items():
                    T[k].append(v)
                T["probs"].append(p0); T["q"].append(q[0]); T["game"].append(gi)
                T["move"].append(-2 if mv is None else int(mv)); T["sampled_legal"].append(int(g.is_legal(sampled)))
                if mv is None:
                    stop = True
                    break
                g.play_move(int(mv))
        finals.append((board_arr(g.board), g.turn, 1 if g.score() > 0 else -1, len(g.moves)))
    Categorical.sample = orig
    return T, finals


def make_playouts(pi17, pi19):
    check_sampler_shim()
    Tm, Fm = trace_mcts(pi17, n_playouts=6, seed0=500)
    Ts, Fs = trace_selfplay(pi17, pi19, n_games=4, seed0=900)
    out = {}
    for pre, T in (("m_", Tm), ("s_", Ts)):
        for k, v in T.items():
            out[pre + k] = np.stack(v) if isinstance(v[0], np.ndarray) else np.array(v)
    out["m_final_board"] = np.stack([f[0] for f in Fm]); out["m_final_turn"] = np.array([f[1] for f in Fm])
    out["m_final_last"] = np.array([f[2] for f in Fm]); out["m_reward"] = np.array([f[3] for f in Fm])
    out["m_score"] = np.array([f[4] for f in Fm], np.float64)
    out["s_final_board"] = np.stack([f[0] for f in Fs]); out["s_final_turn"] = np.array([f[1] for f in Fs])
    out["s_result"] = np.array([f[2] for f in Fs]); out["s_len"] = np.array([f[3] for f in Fs])
    np.savez_compressed(os.path.join(HERE, "playouts.npz"), **out)
    print("mcts playouts: moves", len(Tm["move"]), "passes", int(np.sum(np.array(Tm["move"]) == -1)),
          "max draws/move", int(np.max(Tm["nq"])), "final turns", [f[1] for f in Fm], "rewards", [f[3] for f in Fm])
    print("selfplay games: moves", len(Ts["move"]), "lens", [f[3] for f in Fs],
          "illegal first samples", int(len(Ts["move"]) - np.sum(Ts["sampled_legal"])), "results", [f[2] for f in Fs])


if __name__ == "__main__":
    rec = make_positions()
    make_rules(rec)
    pi17, pi19 = load_policy("policy_17.pt"), load_policy("policy_19.pt")
    save_weights(pi17, "weights_policy_17.npz")
    save_weights(pi19, "weights_policy_19.npz")
    make_nets(rec, pi17, pi19)
    make_playouts(pi17, pi19)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f"{f}: {os.path.getsize(os.path.join(HERE, f)) / 1e6:.2f} MB")
