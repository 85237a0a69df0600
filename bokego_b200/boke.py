"""Command-line GTP engine on the B200 path, flag-compatible with the reference's boke.py (/root/reference/boke.py:13-45):
    python -m bokego_b200.boke -p policy.pt -v value.pt [-t SECONDS | -r ROLLOUTS]
-t / -r / -p / -v as in the reference (its -r is parsed but never forwarded, boke.py:17,40-44; here it works; -t 0 selects
-r).  -g is accepted and ignored: this engine always runs on the GPU.  --simulate enables playouts to the end of the game
from every leaf (boke.py:24-25, MCTS(no_sim=False)): the leaves of a batch are played out together on the device.  Checkpoints are the reference's: torch files holding {"model_state_dict": ...}
(boke.py:30-38), or the .npz state dicts under tests/golden/."""
import argparse

import numpy as np
import torch

from . import nnet
from .gtp import GTP


def load_state(path):
    if path.endswith(".npz"):
        return {k: torch.from_numpy(v) for k, v in np.load(path).items()}
    ck = torch.load(path, map_location="cpu")
    return ck.get("model_state_dict", ck)


def main(argv=None):
    ap = argparse.ArgumentParser(description="BokeGo GTP engine (B200)")
    ap.add_argument("-t", type=float, default=0.0, help="seconds per move (0: use -r)")
    ap.add_argument("-r", type=int, default=1600, help="rollouts per move")
    ap.add_argument("-p", required=True, help="policy net checkpoint")
    ap.add_argument("-v", required=True, help="value net checkpoint")
    ap.add_argument("-g", action="store_true", help="accepted for compatibility; the GPU is always used")
    ap.add_argument("--simulate", action="store_true", help="enable simulations to end of game (batched on the device)")
    ap.add_argument("--leaf-batch", type=int, default=32, help="descents per evaluation batch")
    a = ap.parse_args(argv)
    dev = torch.device("cuda", torch.cuda.current_device())
    pi, v = nnet.PolicyNet(), nnet.ValueNet()
    pi.load_state_dict(load_state(a.p)); v.load_state_dict(load_state(a.v))
    pi.eval(); v.eval()
    GTP(pi, v, time_lim=a.t, n_rollouts=a.r, leaf_batch=a.leaf_batch, no_sim=not a.simulate, device=dev).start()


if __name__ == "__main__":
    main()
