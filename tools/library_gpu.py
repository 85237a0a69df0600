#!/usr/bin/env python
"""Library-GPU comparison: the SAME nets evaluated through stock PyTorch on the B200 (cuDNN / cuBLAS library
kernels) -- the number the hand-written kernel has to beat on the same box.  Measurement tooling, not product code and not
part of bench.py's contract: it evaluates oracle/nets.py (the fp32 restatement of nnet.py:19-113) with the parameters moved
to the GPU, in fp32 (TF32 off and on) and under fp16 autocast with channels_last.
    python tools/library_gpu.py [--batch 4096] > gpurun_out/library_gpu.json"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nets as onets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = os.path.join(ROOT, "tests", "golden")
    sd17 = {k: torch.from_numpy(v).to(dev) for k, v in np.load(os.path.join(g, "weights_policy_17.npz")).items()}
    sdv = {k: torch.from_numpy(v).to(dev) for k, v in np.load(os.path.join(g, "weights_policy_19.npz")).items()}
    sdv.update({k: v.to(dev) for k, v in onets.standin_value_head(1234).items()})
    P = dict(np.load(os.path.join(g, "positions.npz")))
    idx = np.arange(a.batch) % len(P["feats"])
    x = onets.planes_to_float(P["feats"][idx]).to(dev)
    torch.backends.cudnn.benchmark = True
    out = {"batch": a.batch, "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "runs": []}

    def run(name, fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out["runs"].append({"mode": name, "ms_per_batch": ms, "evals_per_s": a.batch / (1e-3 * ms)})

    def both(xx):
        return onets.policy_probs(sd17, xx), onets.value(sdv, xx)

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    run("fp32 (TF32 off)", lambda: both(x))
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    run("fp32 storage, TF32 tensor cores", lambda: both(x))
    xc = x.contiguous(memory_format=torch.channels_last)

    def amp():
        with torch.autocast("cuda", dtype=torch.float16):
            return both(xc)
    run("fp16 autocast, channels_last", amp)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
