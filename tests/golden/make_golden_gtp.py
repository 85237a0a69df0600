#!/usr/bin/env python
"""Golden GTP transcript: a scripted session through the REFERENCE engine (bokego/gtp.py, unmodified, imported from
/root/reference, CPU nets) for every command whose answer does not depend on search results.  bokego_b200.gtp must give the
same answers byte for byte (tests/test_gtp.py).  Runs only in the build container:
    python tests/golden/make_golden_gtp.py   ->  tests/golden/gtp_transcript.json"""
import json
import os
import sys

import torch

REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

import bokego.go as go            # noqa: E402
import bokego.nnet as nnet        # noqa: E402
from bokego.gtp import GTP        # noqa: E402
from bokego.mcts import Go_MCTS   # noqa: E402

SGF = "/tmp/bokego_gtp_golden.sgf"
SESSION = [
    "protocol_version", "name", "version", "1 name", "known_command play", "known_command fly", "known_command",
    "boardsize 9", "boardsize 19", "boardsize", "frobnicate", "clear_board", "komi 6.5", "komi", "komi abc", "showboard",
    "play b e5", "play w e5", "play w d5", "last_move", "play b z9", "play b", "play purple a1", "showboard",
    "play b c3", "play b g7", "move_history", "last_move", "undo", "undo", "move_history", "showboard", "final_score",
    "play w f5", "play b e4", "play w e6", "play b pass", "play w d4", "showboard", "final_score", "12 last_move",
    f"printsgf {SGF}", "clear_board", "last_move", "final_score", "set_fixed_handicap 1", "set_fixed_handicap 3",
    "showboard", "set_fixed_handicap 2", "play w c3", "move_history", "clear_board", f"loadsgf {SGF} 3", "showboard",
    "move_history", "loadsgf /nonexistent.sgf 1", "loadsgf", "pondering", "pondering off", "help", "list_commands",
    "genmove", "genmove purple", "clear_board", "play b a1", "play w a2", "play b j9", "play w b1", "showboard", "play b a1",
    "final_score", "quit", "name",
]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    pol, val = nnet.PolicyNet(), nnet.ValueNet()
    pol.eval(); val.eval()
    g = GTP(Go_MCTS(), pol, val, no_sim=True, time_lim=0, n_rollouts=10, pondering=False)
    g.running = True
    out = []
    for cmd in SESSION:
        out.append({"cmd": cmd, "out": g.send(cmd)})
    json.dump({"sgf_path": SGF, "session": out}, open(os.path.join(HERE, "gtp_transcript.json"), "w"), indent=1)
    for o in out:
        print(repr(o["cmd"]), "->", repr(o["out"])[:100])


if __name__ == "__main__":
    main()
