"""GTP front-end (bokego_b200.gtp) against a scripted session recorded from the unmodified reference engine
(tests/golden/gtp_transcript.json, produced by tests/golden/make_golden_gtp.py): every reply must match byte for byte.
CPU part: commands that need no search.  GPU part: genmove plays legal moves and keeps the protocol state consistent."""
import json
import os

import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_scripted_session_matches_reference(tmp_path):
    from bokego_b200.gtp import GTP
    t = json.load(open(os.path.join(GOLDEN, "gtp_transcript.json")))
    sgf = str(tmp_path / "session.sgf")
    g = GTP()
    g.running = True
    for i, step in enumerate(t["session"]):
        cmd = step["cmd"].replace(t["sgf_path"], sgf)
        want = step["out"]
        got = g.send(cmd)
        assert got == want, (i, cmd, got, want)


def test_command_set_and_reply_format():
    from bokego_b200.gtp import GTP
    g = GTP()
    assert g.send("name") is None                      # not running yet (gtp.py:113-114)
    g.running = True
    assert g.send("7 protocol_version") == "=7 2\n\n"
    assert g.send("") is None
    assert len(GTP.commands) == 24 and "genmove" in GTP.commands
    assert g.send("analyze b 50").startswith("?")


@pytest.mark.gpu
def test_genmove_plays_legal_moves(sd17, sd_value):
    import torch
    from bokego_b200 import batched as bk, go
    from bokego_b200.gtp import GTP
    dev = torch.device("cuda", 0)
    g = GTP(bk.PackedNet(sd17, dev), bk.PackedNet(sd_value, dev), n_rollouts=200, leaf_batch=16, device=dev)
    g.running = True
    assert g.send("clear_board") == "= \n\n"
    seen = []
    for k in range(6):
        color = "b" if k % 2 == 0 else "w"
        before = g.root
        out = g.send(f"genmove {color}")
        assert out.startswith("= ") and out.endswith("\n\n")
        mv = go.squash(out[2:].strip())
        assert 0 <= mv < 81 and before.is_legal(mv) and mv not in seen
        seen.append(mv)
        assert g.root.turn == k + 1 and g.root.last_move == mv
    assert g.send("move_history") == "= " + "\n".join(go.unsquash(seen)) + "\n\n"
    assert g.send("genmove w").startswith("= ")        # white again: black passes first (gtp.py:205-208)
    assert g.root.turn == 8
    assert g.send("undo") == "= \n\n" and g.root.turn == 7      # back to the position after the inserted pass (gtp.py:232-241)
    assert g.send("undo") == "? cannot undo\n\n"                # only one undo
    assert g.send("reg_genmove w").startswith("= ")
    assert g.send("quit") == "= \n\n" and g.send("name") is None
