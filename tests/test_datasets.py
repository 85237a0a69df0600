"""Value-data CSV (the reference's genvals.py / nnet.process_csv format) through the encoder kernel."""
import numpy as np
import pytest

from oracle import cpu as ocpu

DEC = {1: "X", -1: "O", 0: "."}


def _rows(positions, n):
    pick = [i for i in range(len(positions["board"])) if positions["last"][i] >= 0][:n]
    rows = []
    for i in pick:
        board = "".join(DEC[int(v)] for v in positions["board"][i])
        ko = None if positions["ko"][i] < 0 else int(positions["ko"][i])
        rows.append((board, ko, int(positions["last"][i]), i % 2))
    return pick, rows


def test_csv_round_trip(tmp_path, positions):
    from bokego_b200 import datasets
    pick, rows = _rows(positions, 50)
    p = str(tmp_path / "vals.csv")
    datasets.write_value_csv(p, rows[:30])
    datasets.write_value_csv(p, rows[30:], append=True)
    assert open(p).readline().strip() == "board,last,ko,val"
    bd, ko, last, turn, kind, tgt = datasets.read_value_csv(p)
    assert kind == "val" and np.array_equal(bd, positions["board"][pick]) and np.array_equal(last, positions["last"][pick])
    assert np.array_equal(ko, np.where(positions["ko"][pick] < 0, -1, positions["ko"][pick]))
    assert np.array_equal(tgt, np.arange(len(pick)) * 0 + np.array(pick) % 2)
    # side to move: white iff the last stone is black (nnet.py:377)
    assert all(turn[i] == (1 if bd[i, last[i]] == 1 else 0) for i in range(len(pick)))


@pytest.mark.gpu
def test_process_csv_features_match_oracle(tmp_path, positions):
    from bokego_b200 import datasets
    pick, rows = _rows(positions, 300)
    p = str(tmp_path / "vals.csv")
    datasets.write_value_csv(p, rows)
    out = datasets.process_csv(p, str(tmp_path / "vals.npz"))
    bd, ko, last, turn, _, tgt = datasets.read_value_csv(p)
    want, _, _ = ocpu.features_batch(bd, ko, last, turn)                 # fresh Games, as process_csv builds them
    assert out["features"].dtype == np.int8 and out["features"].shape == (len(pick), 27, 9, 9)
    assert np.array_equal(out["features"].reshape(len(pick), 27, 81), want.astype(np.int8))
    assert np.array_equal(out["targets"].ravel(), np.where(tgt != 0, -1, 1))
    z = np.load(str(tmp_path / "vals.npz"))
    assert np.array_equal(z["features"], out["features"]) and np.array_equal(z["targets"], out["targets"])
