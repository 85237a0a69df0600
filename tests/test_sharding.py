"""Host-side multi-rank logic on CPU: shard arithmetic, and the gather of per-game records over two gloo
ranks.  The records are produced by the oracle's self-play stepping keyed by GLOBAL game ids, so the test also
shows that results do not depend on how games are partitioned."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bokego_b200.playout import gather_records, shard_range
from oracle import cpu as ocpu


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 4096, 65537):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 4, 4)


def _oracle_games(lo, hi, seed, steps=12):
    """self-play flavour stepping with a fixed non-uniform policy, games lo..hi-1 of a global numbering"""
    n = hi - lo
    bd = np.zeros((n, 81), np.int8); ko = np.full(n, -1, np.int16); last = np.full(n, -2, np.int16)
    turn = np.zeros(n, np.int16); done = np.zeros(n, np.uint8)
    probs = np.tile((np.arange(81, dtype=np.float32) % 7 + 1.0), (n, 1))
    probs /= probs.sum(1, keepdims=True)
    moves = np.zeros((n, steps), np.int16)
    for k in range(steps):
        moves[:, k] = ocpu.step_batch(bd, ko, last, turn, None, done, probs, 1, 70, seed=seed, game0=lo)
    sc = ocpu.score_batch(bd)
    head = np.stack([turn, np.where(sc > 0, 1, -1).astype(np.int16), np.round(sc * 2).astype(np.int16)], 1)
    return np.concatenate([head, moves], 1).astype(np.int16)


def _worker(rank, world, port, n_total, seed, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    local = torch.from_numpy(_oracle_games(lo, hi, seed))
    full = gather_records(local, n_total, rank, world)
    np.save(os.path.join(out_dir, f"full_{rank}.npy"), full.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gather_equals_single_rank(tmp_path):
    n_total, seed, world = 37, 4242, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, n_total, seed, str(tmp_path)), nprocs=world, join=True)
    single = _oracle_games(0, n_total, seed)
    for r in range(world):
        got = np.load(tmp_path / f"full_{r}.npy")
        assert got.shape == single.shape and np.array_equal(got, single)
    assert len({tuple(row[3:]) for row in single}) > 5      # the games actually differ from one another


def test_gather_is_identity_for_one_rank():
    t = torch.arange(12, dtype=torch.int16).reshape(4, 3)
    assert gather_records(t, 4, 0, 1) is t


def test_record_round_trips_through_sgf(tmp_path):
    """a game record written with the reference's SGF layout reads back move for move"""
    from bokego_b200 import go
    from bokego_b200.playout import record_to_sgf
    rec = _oracle_games(3, 4, seed=99, steps=20)[0]
    path = str(tmp_path / "g.sgf")
    sgf = record_to_sgf(rec, path, B="policy_17", W="policy_19")
    moves = [int(m) for m in rec[3:] if m >= -1]
    assert go.get_moves(path) == moves and len(moves) == 20
    assert "PB[policy_17]PW[policy_19]" in sgf and ("RE[B+" in sgf or "RE[W+" in sgf)
    g = go.Game(sgf=path)                       # like the reference, Game(sgf=...) loads the move list ...
    assert g.moves == moves and len(g) == 20
    for m in g.moves:                           # ... and the moves replay legally
        g.play_move(m)
    assert g.turn == 20


# ---- REINFORCE row: the exchange step of data-parallel training (gather of per-call statistics, sum of gradients) ----------
def _train_worker(rank, world, port, n_games, out_dir):
    from bokego_b200 import reinforce as rf
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_games, rank, world)
    full = torch.arange(3 * n_games * 5, dtype=torch.float32).reshape(3, n_games, 5)       # [step][game][...]
    got = rf.gather_games(full[:, lo:hi].contiguous(), n_games)
    assert rf._world() == world and rf._rank() == rank
    g = torch.full((7,), float(rank + 1))
    dist.all_reduce(g)                                                                      # what reinforce_step does to the gradients
    np.save(os.path.join(out_dir, f"games_{rank}.npy"), got.numpy())
    np.save(os.path.join(out_dir, f"grads_{rank}.npy"), g.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("n_games", [4, 5, 1])
def test_gather_games_two_ranks(tmp_path, n_games):
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_train_worker, args=(world, port, n_games, str(tmp_path)), nprocs=world, join=True)
    full = np.arange(3 * n_games * 5, dtype=np.float32).reshape(3, n_games, 5)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"games_{r}.npy"), full)                  # uneven and empty shards included
        assert np.array_equal(np.load(tmp_path / f"grads_{r}.npy"), np.full(7, 3.0, np.float32))


def test_gather_games_single_rank_is_identity():
    from bokego_b200 import reinforce as rf
    t = torch.zeros(2, 3, 4)
    assert rf.gather_games(t, 3) is t and rf._world() == 1 and rf._rank() == 0
