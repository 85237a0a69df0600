"""Value-net training data on the B200 path: the reference's position CSV -> feature tensors.

The reference generates `(board, ko, last_move, val)` rows with bin/genvals.py (/root/reference/bin/genvals.py:42-45, 73-80:
header line `board,last,ko,val` although the fields are written in the order board, ko, last, val) and turns them into network
inputs with nnet.process_csv (/root/reference/bokego/nnet.py:366-383): every row becomes a fresh go.Game whose side to move is
derived from the colour of the last stone, `features(game)` is stored as int8 (27,9,9) and the target is -1 / +1.  Here the
whole file is encoded by one bk_encode launch.  (The augmentation helpers next to it in the reference, refl / rot, are
undefined there -- genvals.py:79-80 -- and are not reproduced.)"""
import csv

import numpy as np
import torch

from . import go
from .batched import NONE, Positions, features_batch

HEADER = "board,last,ko,val"       # the reference's header; the fields below it are board, ko, last, val


def write_value_csv(path, rows, append=False):
    """rows of (board str, ko int|None, last int, val int) in the reference's layout (genvals.py:42-45)"""
    with open(path, "a+" if append else "w") as f:
        if not append:
            f.write(HEADER + "\n")
        for board, ko, last, val in rows:
            f.write(",".join([board, str(ko), str(last), str(val)]) + "\n")


def read_value_csv(path):
    """-> (boards int8 [N,81], ko int16 [N], last int16 [N], turn int16 [N], target column name, target values)"""
    lut = {go.BLACK: 1, go.WHITE: -1, go.EMPTY: 0}
    with open(path) as f:
        rd = csv.reader(f)
        cols = next(rd)
        rows = [r for r in rd if r]
    n = len(rows)
    bd = np.zeros((n, 81), np.int8); ko = np.full(n, -1, np.int16); last = np.zeros(n, np.int16); tgt = np.zeros(n, np.int64)
    for i, (board, k, l, t) in enumerate(rows):
        bd[i] = [lut[c] for c in board]
        ko[i] = -1 if k.strip() == "None" else int(k)
        last[i] = NONE if l.strip() == "None" else int(l)
        tgt[i] = int(t)
    # nnet.py:377: black to move (turn 0) unless the last stone is black
    turn = np.array([1 if (last[i] >= 0 and bd[i, last[i]] == 1) else 0 for i in range(n)], np.int16)
    return bd, ko, last, turn, cols[-1], tgt


def process_csv(path, npz_name=None, device="cuda"):
    """nnet.process_csv: CSV of positions -> {"features": int8 [N,27,9,9], "targets": int8 [N,1]}; written with
    np.savez_compressed when npz_name is given.  Targets: -1 if val else +1 for a `val` file, the move for a `move` file."""
    bd, ko, last, turn, kind, tgt = read_value_csv(path)
    dev = torch.device(device)
    if len(bd):
        pos = Positions.from_numpy(bd, ko, last, turn, dev)
        fts = features_batch(pos, fresh_libs=True, want=("u8",))["u8"].cpu().numpy().astype(np.int8).reshape(-1, 27, 9, 9)
    else:
        fts = np.zeros((0, 27, 9, 9), np.int8)
    targets = np.where(tgt != 0, -1, 1).astype(np.int8) if kind == "val" else tgt.astype(np.int8)
    out = {"features": fts, "targets": targets.reshape(-1, 1)}
    if npz_name:
        np.savez_compressed(npz_name, **out)
    return out
