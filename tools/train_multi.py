#!/usr/bin/env python
"""Data-parallel REINFORCE over R GPUs (torchrun, one rank per GPU) against the single-GPU run: same games (the random stream is
keyed by the global game id), same win counts, parameters equal up to the order of the gradient sum.
    torchrun --standalone --nproc-per-node 2 tools/train_multi.py        (also run with one process to write the baseline)"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import nnet, reinforce as rf  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
g = os.path.join(ROOT, "tests", "golden")
sd17 = dict(np.load(os.path.join(g, "weights_policy_17.npz")))
sd19 = dict(np.load(os.path.join(g, "weights_policy_19.npz")))
out = {}
for acc, bs in (("batch", 6), ("reference", 5)):
    pi, opp = nnet.PolicyNet(), nnet.PolicyNet()
    pi.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd17.items()})
    opp.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd19.items()})
    pi.to(dev).train()
    opp.to(dev).eval()
    opt = torch.optim.AdamW(pi.parameters(), lr=1e-5)
    stats = []
    tr = rf.reinforce(pi, opp, opt, "white", n_itrs=2, bs=bs, device=dev, stats=stats, seed=3, accumulate=acc)
    torch.cuda.synchronize()
    out[acc] = {"wins": stats, "grads": tr.grads.cpu().numpy(), "params": tr.params.cpu().numpy(), "running": tr.running.cpu().numpy()}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
if world == 1:
    np.savez(os.path.join(ROOT, "gpurun_out", "train_multi_base.npz"),
             **{f"{a}/{k}": np.asarray(v) for a, d in out.items() for k, v in d.items()})
    print("baseline written", {a: d["wins"] for a, d in out.items()})
else:
    base = np.load(os.path.join(ROOT, "gpurun_out", "train_multi_base.npz"))
    rep = {"world": world, "rank": rank}
    for a, d in out.items():
        gs = float(np.abs(base[f"{a}/grads"]).max())
        rep[a] = {"wins_equal": list(base[f"{a}/wins"]) == d["wins"],
                  "grad_err_rel": float(np.abs(d["grads"] - base[f"{a}/grads"]).max() / gs),
                  "param_err": float(np.abs(d["params"] - base[f"{a}/params"]).max()),
                  "running_err": float(np.abs(d["running"] - base[f"{a}/running"]).max())}
        # parameters: an entry whose gradient is at round-off level (the conv biases in front of a BatchNorm) gets a step of +-lr from
        # AdamW's normalisation whatever its sign, and the order of the gradient sum decides the sign: up to 2 lr per iteration
        assert rep[a]["wins_equal"] and rep[a]["grad_err_rel"] < 1e-3 and rep[a]["param_err"] <= 2 * 2.1e-5 and rep[a]["running_err"] < 1e-3, rep
    # replicas identical across ranks
    p = tr.params.clone()
    dist.broadcast(p, 0)
    rep["replicas_identical"] = bool(torch.equal(p, tr.params))
    assert rep["replicas_identical"]
    print(json.dumps(rep))
    dist.destroy_process_group()
