#!/usr/bin/env python
"""bench.py -- BokeGo hot-path benchmark on B200 (contract in the task statement).

Metric (BASELINE.json): policy+value position evaluations per second at batch 4096 -- one "step" is one
pass of the hot path over one batch of synthetic legal 9x9 positions:
    boards -> bk_encode (27 feature planes) -> bk_forward (PolicyNet + ValueNet, softmax / tanh fused).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B]        our arm (CUDA kernels via the C ABI)
    python bench.py --impl reference ...                                   the CPU arm: the unmodified reference from baseline/_ref
                                                                           on the host cores (the oracle port if it is not installed)

N > 1 is launched by the driver as torchrun, one rank per GPU; positions are independent, so every rank
evaluates its own batch (weak scaling, no collective on the data path) and rank 0 reports the aggregate.
Only the cpu_baseline / --impl reference legs touch oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_VALID = 266_838_272     # policy+value eval, zero-padding MACs excluded -- primary
FLOP_DENSE = 314_700_032     # what nn.Conv2d executes with zero padding -- secondary


WORKLOAD = "batched PolicyNet+ValueNet forward incl. feature encoding, batch 4096 synthetic legal 9x9 positions (BASELINE configs[1])"


def bench_config(batch):
    """the `config` object of BOTH arms (identical keys and values, so the driver can compare the lines)"""
    return {"workload": WORKLOAD, "batch_per_gpu": batch, "weights": "policy_17 + stand-in ValueNet (policy_19 trunk, seeded head)",
            "positions": "seeded uniformly-random legal play, depth ~ U{0..70}, evaluated as fresh positions",
            "l2": "GPU arm: L2 flushed between timed iterations (256 MiB memset); CPU arm: not applicable"}


def reference_root():
    """the unmodified reference installed by tools/install_reference.sh (git-ignored, travels to the GPU box), or None"""
    r = os.path.join(ROOT, "baseline", "_ref")
    return r if os.path.isfile(os.path.join(r, "bokego", "nnet.py")) else None


def import_reference():
    """the reference's own modules (go, nnet, mcts, gtp) from baseline/_ref -- used by the CPU legs only"""
    import importlib
    r = reference_root()
    if r is None:
        return None
    for k in [k for k in sys.modules if k == "bokego" or k.startswith("bokego.")]:
        del sys.modules[k]
    if r not in sys.path:
        sys.path.insert(0, r)
    mods = tuple(importlib.import_module("bokego." + m) for m in ("go", "nnet", "mcts", "gtp"))
    assert os.path.abspath(mods[1].__file__).startswith(os.path.abspath(r)), "bokego.nnet is not the reference's"
    return mods


def reference_nets(nnet, sd17, sdv):
    """the reference's nn.Modules with the bench weights (CPU, eval)"""
    t = lambda sd: {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
    pi, v = nnet.PolicyNet(), nnet.ValueNet()
    pi.load_state_dict(t(sd17))
    v.load_state_dict(t(sdv))
    return pi.eval(), v.eval()


def load_nets():
    g = os.path.join(ROOT, "tests", "golden")
    sd17 = dict(np.load(os.path.join(g, "weights_policy_17.npz")))
    sd19 = dict(np.load(os.path.join(g, "weights_policy_19.npz")))
    sdv = dict(sd19)
    # seeded stand-in value head (the reference ships no value_1.pt): the same numbers as
    # oracle.nets.standin_value_head(1234), kept as a fixture so that this arm does not import oracle/
    sdv.update(dict(np.load(os.path.join(g, "weights_value_head_standin.npz"))))
    return sd17, sd19, sdv


def ncu_traffic(kernel, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this kernel
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py); None when no capture exists for this batch size"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    e = json.load(open(p)).get(kernel)
    return e["dram_bytes_per_launch"] if e and e.get("batch") == batch else None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops"], d["hbm_gbs"], "measured (MEASURED_PEAKS.json, burst)"
    return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled in a background thread DURING the timed region (NVML in-process, every
    2 ms; nvidia-smi as a fallback when the bindings are missing)"""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons, self.stop, self.n = index, [], [], set(), False, 0
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical(index))
        except Exception:  # noqa: BLE001
            self.nv = None
        self.t = threading.Thread(target=self.run, daemon=True)

    @staticmethod
    def _physical(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if index < len(ids) and ids[index].strip().isdigit():
                return int(ids[index])
        return index

    def run(self):
        while not self.stop:
            try:
                if self.nv is not None:
                    nv = self.nv
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                                      (0x4, "sw_power_cap")):
                        if r & bit:
                            self.reasons.add(name)
                    self.n += 1
                    time.sleep(0.002)
                    continue
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    c = [x.strip() for x in o.split(",")]
                    self.sm.append(float(c[0])); self.mx.append(float(c[1])); self.n += 1
                    for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                        if c[2 + i].lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_min_mhz": min(self.sm) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons), "samples": self.n,
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


def synth_positions(bk, dev, n, seed):
    """seeded uniformly-random legal play, depth ~ U{0..70}, generated by the stepping kernel"""
    rng = np.random.default_rng(seed)
    depth = torch.from_numpy(rng.integers(0, 71, n).astype(np.int16)).to(dev)
    pos = bk.Positions.empty(n, dev)
    uni = torch.full((n, 81), 1.0 / 81, dtype=torch.float32, device=dev)
    for _ in range(71):
        pos.done.copy_(((pos.turn >= depth) | (pos.last == -1)).to(torch.uint8))
        bk.playout_step(pos, uni, bk.MODE_MCTS, 1000, seed=seed, game0=0)
    pos.done.zero_()
    pos.libs = None          # evaluate as fresh positions (MCTS leaves are scored once each)
    torch.cuda.synchronize()
    return pos


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (C feature encoder on all cores + fp32 torch forward with all threads)
# ---------------------------------------------------------------------------------------------------------
def cpu_eval(bd, ko, last, turn, sd17, sdv):
    from oracle import cpu as ocpu
    from oracle import nets as onets
    f, _, _ = ocpu.features_batch(bd, ko, last, turn, None)
    x = onets.planes_to_float(f)
    p = onets.policy_probs(sd17, x)
    v = onets.value(sdv, x)
    return p, v


def cpu_positions(n, seed):
    """host-side generator of the same kind of positions (random legal play through the oracle)"""
    from oracle import cpu as ocpu
    rng = np.random.default_rng(seed)
    depth = rng.integers(0, 71, n).astype(np.int16)
    bd = np.zeros((n, 81), np.int8); ko = np.full(n, -1, np.int16); last = np.full(n, -2, np.int16)
    turn = np.zeros(n, np.int16); done = np.zeros(n, np.uint8)
    uni = np.full((n, 81), 1.0 / 81, np.float32)
    for _ in range(71):
        done[:] = ((turn >= depth) | (last == -1)).astype(np.uint8)
        ocpu.step_batch(bd, ko, last, turn, None, done, uni, 0, 1000, seed=seed, game0=0)
    return bd, ko, last, turn


def cpu_rate(sample, sd17, sdv, seed=1, repeats=1):
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    sd17t = {k: torch.from_numpy(np.asarray(v)) for k, v in sd17.items()}
    sdvt = {k: torch.from_numpy(np.asarray(v)) for k, v in sdv.items()}
    bd, ko, last, turn = cpu_positions(sample, seed)
    cpu_eval(bd[:64], ko[:64], last[:64], turn[:64], sd17t, sdvt)     # warm-up
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        cpu_eval(bd, ko, last, turn, sd17t, sdvt)
        ts.append(time.perf_counter() - t0)
    return sample / float(np.mean(ts)), cores, ts


def cpu_selfplay_rate(n_games, sd17, sd19, seed=1):
    """CPU arm of the self-play line: the same 72-move games (policy_17 vs policy_19) through the oracle port --
    C feature encoder + fp32 torch CPU policy forward + C stepping -- all games of the sample in lock step on all cores"""
    from oracle import cpu as ocpu
    from oracle import nets as onets
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    nets = [{k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()} for sd in (sd17, sd19)]
    bd = np.zeros((n_games, 81), np.int8); ko = np.full(n_games, -1, np.int16); last = np.full(n_games, -2, np.int16)
    turn = np.zeros(n_games, np.int16); done = np.zeros(n_games, np.uint8); libs = None
    t0 = time.perf_counter()
    for k in range(72):
        f, _, libs = ocpu.features_batch(bd, ko, last, turn, libs)
        probs = onets.policy_probs(nets[k % 2], onets.planes_to_float(f)).numpy()
        ocpu.step_batch(bd, ko, last, turn, libs, done, probs, 1, 70, seed=seed, game0=0)
    ocpu.score_batch(bd)
    dt = time.perf_counter() - t0
    return n_games / dt, cores, dt


def cpu_simulate_rate(n_boards, sd17, seed=2):
    """CPU arm of the --simulate line: MCTS-flavour playouts (Go_MCTS.find_random_child to the end, mcts.py:195-206) through the
    oracle port -- C feature encoder + fp32 torch CPU policy forward + C stepping -- all boards of the sample in lock step"""
    from oracle import cpu as ocpu
    from oracle import nets as onets
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    net = {k: torch.from_numpy(np.asarray(v)) for k, v in sd17.items()}
    bd = np.zeros((n_boards, 81), np.int8); ko = np.full(n_boards, -1, np.int16); last = np.full(n_boards, -2, np.int16)
    turn = np.zeros(n_boards, np.int16); done = np.zeros(n_boards, np.uint8); libs = None
    t0 = time.perf_counter()
    for k in range(81):
        f, _, libs = ocpu.features_batch(bd, ko, last, turn, libs)
        probs = onets.policy_probs(net, onets.planes_to_float(f)).numpy()
        ocpu.step_batch(bd, ko, last, turn, libs, done, probs, 0, 80, seed=seed, game0=0)
        if done.all():
            break
    ocpu.score_batch(bd)
    dt = time.perf_counter() - t0
    return n_boards / dt, cores, dt


def cpu_genmove_reference(sd17, sdv, n_rollouts):
    """BASELINE configs[0] / the CPU side of configs[2]: the reference's own engine on the host CPU, unmodified --
    GTP(Go_MCTS(), pi, v, no_sim=True, time_lim=0, n_rollouts=n).send("genmove b") (gtp.py:344-366; boke.py's `-r` is never
    forwarded, SURVEY 3.1, so the engine is constructed directly).  One thread, as the reference runs it."""
    mods = import_reference()
    if mods is None:
        return None
    go, nnet, mcts, gtp = mods
    torch.set_num_threads(1)
    pi, v = reference_nets(nnet, sd17, sdv)
    for c in (mcts.MCTS._val_cache, mcts.MCTS._dist_cache, mcts.MCTS._fts_cache):
        c.clear()
    with torch.no_grad():
        t0 = time.perf_counter()
        g = gtp.GTP(mcts.Go_MCTS(), pi, v, no_sim=True, time_lim=0, n_rollouts=n_rollouts, pondering=False)
        g.running = True
        out = g.send("genmove b")
        dt = time.perf_counter() - t0
    return {"seconds": dt, "playouts_per_s": n_rollouts / dt, "move": out.strip("= \n"), "value_evals": len(mcts.MCTS._val_cache),
            "policy_evals": len(mcts.MCTS._dist_cache), "cores": 1, "kind": "reference",
            "sample": f"one genmove, {n_rollouts} rollouts from the empty board, nets on the CPU, torch threads 1"}


def gpu_genmove_reference_callers(sd17, sdv, dev, n_rollouts, batched):
    """The same engine -- the reference's gtp.py / mcts.py, unmodified, from baseline/_ref -- over the B200 mirror
    (bokego_b200.dropin), nets on the GPU; `batched`: children of every expanded node evaluated in one launch pair through
    the class-level caches (dropin.batch_expansions)."""
    from bokego_b200 import _lib, dropin, nnet
    if dropin.find_reference() is None:
        return None
    mcts, gtp, _ = dropin.reference_callers()
    try:
        t = lambda sd: {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
        pi, v = nnet.PolicyNet(), nnet.ValueNet()
        pi.load_state_dict(t(sd17)); v.load_state_dict(t(sdv))
        pi.eval().to(dev); v.eval().to(dev)
        dropin.batch_expansions(mcts.MCTS, batched)
        res = None
        for rep in range(2):                     # the first pass warms the packed blobs and the kernels
            for c in (mcts.MCTS._val_cache, mcts.MCTS._dist_cache, mcts.MCTS._fts_cache):
                c.clear()
            n0 = _lib.launch_count
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            g = gtp.GTP(mcts.Go_MCTS(), pi, v, no_sim=True, time_lim=0, n_rollouts=n_rollouts, pondering=False, device=dev)
            g.running = True
            out = g.send("genmove b")
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            res = {"seconds": dt, "playouts_per_s": n_rollouts / dt, "move": out.strip("= \n"), "kernel_launches": _lib.launch_count - n0,
                   "value_evals": len(mcts.MCTS._val_cache), "batched_expansions": bool(batched)}
        return res
    finally:
        dropin.batch_expansions(mcts.MCTS, False)
        dropin.uninstall()


FLOP_TRAIN = 2 * (2 * 66_706_944 + 6 * 10_240_000 + 10_368)   # forward + weight gradient + data gradient, valid taps, per position


def reinforce_leg(dev, sd17, sd19, with_cpu):
    """REINFORCE row (bin/selfplay.py:59-122): the training step on the 576 positions of a reference-sized
    batch (bs = 16 games x 36 moves of the training colour), and whole iterations of the loop (self-play + step)."""
    from bokego_b200 import nnet, reinforce as rf
    calls = np.load(os.path.join(ROOT, "tests", "golden", "reinforce.npz"))["black3/calls"]
    P = 576
    planes = torch.from_numpy(np.ascontiguousarray(calls[np.arange(P) % len(calls)])).to(dev)
    moves = torch.randint(0, 81, (P,), device=dev).to(torch.int16)
    coef = torch.full((P,), 1.0 / 16, device=dev)
    out = {"positions": P, "flop_per_position": FLOP_TRAIN}
    tf_peak = peaks()[0] / 2            # TF32 runs at half the 16-bit tensor rate
    for prec, name in ((rf.PREC_TC_3XTF32, "3xtf32 (tcgen05; default, fp32-grade)"), (rf.PREC_TC_TF32, "tf32 (tcgen05)"),
                       (rf.PREC_3XTF32, "3xtf32_mma_sync (warp-level MMA)"), (rf.PREC_TF32, "tf32_mma_sync (warp-level MMA)")):
        tr = rf.PolicyTrainer(sd17, dev, prec=prec)
        for _ in range(3):
            rf.reinforce_step(tr, planes, moves, coef)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        a.record()
        for _ in range(n):
            rf.reinforce_step(tr, planes, moves, coef)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        ach = FLOP_TRAIN * P / (1e-3 * ms) / 1e12
        out[name.split(" ")[0]] = {"step_ms": ms, "positions_per_s": P / (1e-3 * ms), "achieved_tflops": ach,
                                   "frac_of_tf32_peak": ach / tf_peak, "precision": name}
    out["peak"] = {"tflops": tf_peak, "source": "half of the measured 16-bit dense peak (TF32 operands)",
                   "note": "clock stamps of the 3x3 kernel (profiles/r02u_train_conv3.md): a kind::tf32 MMA of 128x128x8 takes ~104 cycles, "
                           "~745 TFLOP/s at 1.92 GHz with M = N = 128 cta_group::1 MMAs; a 3xTF32 product issues three times its useful work, so "
                           "~248 TFLOP/s of useful work is the ceiling of the default precision with these MMAs"}
    # the same step on a batch that fills the GPU (2,048 positions = one forward / backward chunk), default precision
    P2 = 2048
    planes2 = torch.from_numpy(np.ascontiguousarray(calls[np.arange(P2) % len(calls)])).to(dev)
    moves2 = torch.randint(0, 81, (P2,), device=dev).to(torch.int16)
    coef2 = torch.full((P2,), 1.0 / 16, device=dev)
    tr = rf.PolicyTrainer(sd17, dev)
    for _ in range(2):
        rf.reinforce_step(tr, planes2, moves2, coef2)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        rf.reinforce_step(tr, planes2, moves2, coef2)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    ach = FLOP_TRAIN * P2 / (1e-3 * ms) / 1e12
    out["at_2048_positions"] = {"step_ms": ms, "positions_per_s": P2 / (1e-3 * ms), "achieved_tflops": ach, "frac_of_tf32_peak": ach / tf_peak,
                                "precision": "3xtf32 (tcgen05; default)"}
    del tr, planes2
    pi, opp = nnet.PolicyNet(), nnet.PolicyNet()
    pi.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd17.items()})
    opp.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd19.items()})
    pi.to(dev).train()
    opp.to(dev).eval()
    opt = torch.optim.AdamW(pi.parameters(), lr=1e-5)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):        # reinforce() prints the win rate like the reference; stdout carries the JSON line only
        rf.reinforce(pi, opp, opt, "black", n_itrs=2, bs=16, device=dev, stats=[])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 10
        rf.reinforce(pi, opp, opt, "black", n_itrs=n, bs=16, device=dev, stats=[], seed=100)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    out["iteration"] = {"bs": 16, "seconds": dt, "games_per_s": 16 / dt, "iterations_per_s": 1 / dt,
                        "what": "reinforce(): 16 self-play games in lock step (pi in train mode vs policy_19), running statistics, "
                                "backward of the last game (the reference's per-game loss reset), AdamW; host wall clock"}
    if with_cpu:
        from oracle import train as ot
        cores = len(os.sched_getaffinity(0))
        torch.set_num_threads(cores)
        n_cpu = 72
        x = calls[:n_cpu].astype(np.float32)
        mv, cf = np.arange(n_cpu) % 81, np.full(n_cpu, 1.0 / 16, np.float32)
        ot.reinforce_grads(sd17, x[:8], mv[:8], cf[:8])
        t0 = time.perf_counter()
        _, grads, _ = ot.reinforce_grads(sd17, x, mv, cf)
        for k, g in grads.items():
            ot.adamw_step(sd17[k], g.numpy(), np.zeros_like(sd17[k]), np.zeros_like(sd17[k]), 1)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "positions/s", "cores": cores, "kind": "port",
                               "sample": f"{n_cpu} positions: torch CPU autograd forward + backward + AdamW restatement ({dt:.2f} s)"}
    return out


_REF_WORKER = {}


def _ref_worker_init(root):
    """pool worker: import the reference once (CPU only)"""
    sys.path.insert(0, root)
    import bokego.go as go, bokego.nnet as nnet      # noqa: E401
    torch.set_num_threads(1)
    _REF_WORKER["go"], _REF_WORKER["nnet"] = go, nnet


def _ref_worker_features(chunk):
    """nnet.features (/root/reference/bokego/nnet.py:182-262) of a chunk of positions, each a fresh go.Game"""
    go, nnet = _REF_WORKER["go"], _REF_WORKER["nnet"]
    dec = {1: go.BLACK, -1: go.WHITE, 0: go.EMPTY}
    out = []
    for bd, ko, last, turn in chunk:
        g = go.Game("".join(dec[int(v)] for v in bd), None if ko < 0 else int(ko), None if last == -2 else int(last), int(turn))
        out.append(nnet.features(g).numpy())
    return np.stack(out)


class ReferenceEvaluator:
    """The reference's own CPU implementation of the hot path, unmodified: nnet.features per position (one process per host
    core) and PolicyNet / ValueNet fp32 forward over the batch with all torch threads (8d (ii))."""

    def __init__(self, sd17, sdv):
        import multiprocessing as mp
        self.go, self.nnet, _, _ = import_reference()
        self.cores = len(os.sched_getaffinity(0))
        self.pi, self.v = reference_nets(self.nnet, sd17, sdv)
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_ref_worker_init, initargs=(reference_root(),))

    def evaluate(self, bd, ko, last, turn):
        n = len(bd)
        step = max(1, (n + 4 * self.cores - 1) // (4 * self.cores))
        chunks = [[(bd[i], int(ko[i]), int(last[i]), int(turn[i])) for i in range(lo, min(n, lo + step))] for lo in range(0, n, step)]
        x = torch.from_numpy(np.concatenate(self.pool.map(_ref_worker_features, chunks)))
        torch.set_num_threads(self.cores)
        with torch.no_grad():
            p = self.nnet.SOFT(self.pi(x))
            v = self.v(x).reshape(-1)
        return p, v

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args, rank, world):
    """The CPU arm: the reference's own implementation on the host cores when baseline/_ref holds it (kind "reference"),
    else the oracle port (kind "port").  Every step evaluates a bounded sample of the 4096-position batch."""
    if rank != 0:
        return
    sd17, _, sdv = load_nets()
    cores = len(os.sched_getaffinity(0))
    ref = ReferenceEvaluator(sd17, sdv) if reference_root() is not None else None
    if ref is not None:
        kind = "reference"
        how = f"unmodified reference from baseline/_ref: nnet.features in {cores} processes + PolicyNet/ValueNet fp32 torch forward, {cores} threads"
        evaluate = ref.evaluate
    else:
        kind = "port"
        how = f"C feature oracle (OpenMP) + fp32 torch CPU forward, {cores} threads"
        torch.set_num_threads(cores)
        sd17t = {k: torch.from_numpy(np.asarray(v)) for k, v in sd17.items()}
        sdvt = {k: torch.from_numpy(np.asarray(v)) for k, v in sdv.items()}
        evaluate = lambda *a: cpu_eval(*a, sd17t, sdvt)
    # bounded sample: size it so that W + K steps end within about two minutes
    bd, ko, last, turn = cpu_positions(args.batch, 1)
    evaluate(bd[:64], ko[:64], last[:64], turn[:64])
    t0 = time.perf_counter()
    evaluate(bd[:256], ko[:256], last[:256], turn[:256])
    probe = 256 / (time.perf_counter() - t0)
    budget_s = 120.0 / max(1, args.steps + args.warmup)
    sample = int(max(256, min(args.batch, probe * budget_s) // 256 * 256))
    bd, ko, last, turn = bd[:sample], ko[:sample], last[:sample], turn[:sample]
    for _ in range(args.warmup):
        evaluate(bd, ko, last, turn)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        evaluate(bd, ko, last, turn)
    el = time.perf_counter() - t0
    if ref is not None:
        ref.close()
    val = sample * args.steps / el
    line = {"impl": "reference", "metric": "policy+value evals/sec", "value": val, "unit": "evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": bench_config(args.batch),
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": kind,
                             "sample": f"{sample} of the {args.batch} positions per step x {args.steps} steps; {how}"},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    from bokego_b200 import _lib, batched as bk
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    sd17, sd19, sdv = load_nets()
    pol, val = bk.PackedNet(sd17, dev), bk.PackedNet(sdv, dev)
    pos = synth_positions(bk, dev, B, seed=1 + rank)
    L = _lib.lib()
    feats = {"conv": torch.empty(L.bk_feats_conv_bytes(B), dtype=torch.uint8, device=dev),
             "legal": torch.empty(B, 81, dtype=torch.uint8, device=dev)}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    outs = {"legal": feats["legal"]}
    logits_o, probs_o, value_o = (torch.empty(B, 81, device=dev), torch.empty(B, 81, device=dev), torch.empty(B, device=dev))

    def step():
        bk.features_batch(pos, fresh_libs=True, want=("conv", "legal"), out=feats)
        return bk.policy_value_batch(feats["conv"], B, pol, val, want_logits=True)

    def step_one_launch():
        # the same step as ONE launch: bk_forward_positions encodes the planes of every item on chip (the next item's while the
        # tensor pipe works on the current one) and runs both nets; timed beside the headline (it wins per call and at
        # small batches; the two-launch form keeps the better conv-kernel time, and pipelines better end to end)
        return bk.evaluate_positions(pos, pol, val, fresh_libs=True, want_logits=True, want=("legal",), out=outs, probs_out=probs_o,
                                     value_out=value_o)

    # end-to-end leg: positions start in pinned HOST memory and results end there (bk.HostEvaluator is the call a
    # CPU-side search makes): every step copies its inputs in and its probabilities / values out
    def host_eval(depth):
        ev = bk.HostEvaluator(B, pol, val, dev, depth=depth)
        for i in range(depth):
            for h, t in zip(ev.slot(i)["h"], (pos.boards, pos.ko, pos.last, pos.turn)):
                h.copy_(t.cpu())
        return ev
    ev_sync, ev_pipe = host_eval(1), host_eval(3)

    def step_e2e():
        ev_sync.run()

    def timed(fn, n, warm, flush_l2=True):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in ev:
            if flush_l2:
                flush.zero_()                  # evict L2 between timed iterations (outside the timed span)
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        ms = [a.elapsed_time(b) for a, b in ev]
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), ms

    idx = torch.cuda.current_device()
    n0 = _lib.launch_count
    with ClockSampler(idx) as cs:
        tot_ms, ms = timed(step, args.steps, args.warmup)
        launches = _lib.launch_count - n0
        # the dominant kernel alone (roofline)
        def fwd_only():
            bk.policy_value_batch(feats["conv"], B, pol, val, want_logits=True)
        k_tot, k_ms = timed(fwd_only, args.steps, 1)
        t1_tot, t1_ms = timed(step_one_launch, args.steps, 1)
        def enc_only():
            bk.features_batch(pos, fresh_libs=True, want=("conv", "legal"), out=feats)
        e_tot, e_ms = timed(enc_only, args.steps, 1)
        e2e_tot, e2e_ms = timed(step_e2e, args.steps, max(1, args.warmup // 2))
        # the same with rotated staging buffers (copies of one call under the kernels of the next), timed as one region
        for _ in range(3):
            ev_pipe.run()
        ev_pipe.drain()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(args.steps):
            ev_pipe.run()
        ev_pipe.drain()
        p1.record()
        torch.cuda.synchronize()
        pipe_tot = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(pipe_tot, op=dist.ReduceOp.MAX)
        pipe_tot = float(pipe_tot.item())
        # secondary lines of the metric: self-play games/s (args.selfplay_games games sharded over the ranks: strong
        # scaling, gather of the records inside the timed region) and --simulate playouts/s (weak scaling)
        extra = {}
        if not args.no_playouts:
            from bokego_b200 import playout as po
            pol19 = bk.PackedNet(sd19, dev)
            lo, hi = po.shard_range(args.selfplay_games, rank, world)
            sp = po.PlayoutGraph(hi - lo, dev, pol, bk.MODE_SELFPLAY, seed=1, game0=lo, policy_odd=pol19)

            def selfplay_once():
                res = sp.replay()
                po.gather_records(res.records(), args.selfplay_games, rank, world)
            sp_tot, sp_ms = timed(selfplay_once, 3, 1, flush_l2=False)
            # the same as a steady stream of batches (what a self-play worker does): the gather of batch i runs on a side stream
            # under the games of batch i + 1; n_pipe batches timed as ONE region, max over ranks
            n_pipe = 6
            side = torch.cuda.Stream(device=dev)
            keep = []

            def selfplay_stream(n):
                main = torch.cuda.current_stream(dev)
                for _ in range(n):
                    rec = sp.replay().records().clone()          # the graph's record buffer is reused by the next batch
                    done = torch.cuda.Event()
                    done.record(main)
                    side.wait_event(done)
                    with torch.cuda.stream(side):
                        keep.append((rec, po.gather_records(rec, args.selfplay_games, rank, world)))
                    del keep[:-2]
                main.wait_stream(side)
            sp_pipe_tot = None
            if world > 1:                        # (with one GPU there is no gather to overlap)
                selfplay_stream(2)
                sp_pipe_tot, _ = timed(lambda: selfplay_stream(n_pipe), 1, 0, flush_l2=False)
            sim = po.PlayoutGraph(args.simulate_boards, dev, pol, bk.MODE_MCTS, seed=2, game0=rank * args.simulate_boards)
            sim_tot, sim_ms = timed(lambda: sim.replay(), 2, 1, flush_l2=False)
            extra = {"selfplay": {"games": args.selfplay_games, "games_per_s": args.selfplay_games * 3 / (1e-3 * sp_tot),
                                  "ms_per_batch": sp_tot / 3, "moves_per_game": 72, "scaling": "strong",
                                  "nets": "policy_17 (black) vs policy_19 (white)", "gather": "nccl all_gather of records" if world > 1 else "none",
                                  "kernel_launches_per_batch": sp.launches, "boards_per_gpu": hi - lo, "us_per_move": 1e3 * sp_tot / 3 / 72,
                                  "engine": "persistent playout kernel (bk_playout_run: whole games in one launch)" if sp.persistent
                                  else "two launches per move (bk_forward + bk_playout_step_encode), CUDA graph",
                                  "steady_stream": None if sp_pipe_tot is None else {
                                      "games_per_s": args.selfplay_games * n_pipe / (1e-3 * sp_pipe_tot), "ms_per_batch": sp_pipe_tot / n_pipe,
                                      "batches": n_pipe, "how": "batches back to back, the result gather of one batch on a side stream under "
                                                                "the games of the next; one timed region"}},
                     "simulate": {"boards_per_gpu": args.simulate_boards, "playouts_per_s": world * args.simulate_boards * 2 / (1e-3 * sim_tot),
                                  "ms_per_batch": sim_tot / 2, "scaling": "weak", "max_turn": 80}}
            # weak-scaling self-play beside the strong-scaling line: args.selfplay_games games PER GPU
            if world > 1:
                spw = po.PlayoutGraph(args.selfplay_games, dev, pol, bk.MODE_SELFPLAY, seed=1, game0=rank * args.selfplay_games, policy_odd=pol19)
                spw_tot, _ = timed(lambda: spw.replay(), 3, 1, flush_l2=False)
                extra["selfplay"]["weak_scaling"] = {"games_per_gpu": args.selfplay_games, "persistent_kernel": spw.persistent,
                                                     "games_per_s": world * args.selfplay_games * 3 / (1e-3 * spw_tot), "ms_per_batch": spw_tot / 3}
                del spw
            # BASELINE configs[2] and [0]: one genmove from the empty board -- rank 0 only (a host-side search per rank would only
            # make the ranks fight for host cores); the other ranks wait at the barrier below
            if rank == 0:
                from bokego_b200 import mcts as bmcts
                bmcts.MCTS(None, pol, val, device=dev).rollout(10)
                gm = {}
                for name, n_roll, kw in (("reference_parameters", 1600, {"expand_thresh": 100, "leaf_batch": 32}),
                                         ("expand_on_second_visit", 1600, {"expand_thresh": 1, "leaf_batch": 128}),
                                         ("simulate_mode", 1600, {"expand_thresh": 100, "leaf_batch": 64, "no_sim": False}),
                                         ("config0_200_rollouts", 200, {"expand_thresh": 100, "leaf_batch": 1})):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    tree = bmcts.MCTS(None, pol, val, device=dev, **kw)
                    tree.rollout(n_roll)
                    tree.choose()
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    gm[name] = dict(kw, playouts=n_roll, seconds=dt, playouts_per_s=n_roll / dt, net_evals=tree.n_evals,
                                    eval_batches=tree.n_eval_batches, device_playouts=tree.n_playouts)
                extra["mcts_genmove"] = dict(gm, position="empty 9x9 board", timing="host wall clock around the search")
                # the reference's own gtp.py / mcts.py over the mirror on the GPU (row b'), 200 rollouts = BASELINE configs[0]
                rc = {}
                for name, batched in (("one_position_per_call", False), ("batched_expansions", True)):
                    r = gpu_genmove_reference_callers(sd17, sdv, dev, 200, batched)
                    if r is not None:
                        rc[name] = r
                extra["genmove_reference_callers"] = rc if rc else {"unavailable": "no reference install in baseline/_ref"}
            if dist:
                dist.barrier()
        if world == 1 and not args.no_train:
            extra["reinforce"] = reinforce_leg(dev, sd17, sd19, not args.no_cpu)
    clocks = cs.summary()
    launches_timed = launches * args.steps // (args.steps + args.warmup)

    if rank == 0:
        tf_peak, hbm_peak, peak_src = peaks()
        k_s = 1e-3 * float(np.mean(k_ms))
        ach = FLOP_VALID * B / k_s / 1e12
        value = world * B * args.steps / (1e-3 * tot_ms)
        e2e = world * B * args.steps / (1e-3 * e2e_tot)
        line = {
            "metric": "policy+value evals/sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": bench_config(B), "operands": "fp16 operands, fp32 accumulation (tcgen05 kind::f16)",
            "clocks": clocks,
            "e2e": {"value": world * B * args.steps / (1e-3 * pipe_tot), "unit": "evals/s", "h2d_bytes_per_step": ev_pipe.h2d_bytes,
                    "d2h_bytes_per_step": ev_pipe.d2h_bytes,
                    "how": "bk.HostEvaluator(depth=3): pinned host buffers, one H2D + encode + forward + one D2H per step, staging rotated "
                           "so copies and the encoder overlap the previous step's conv kernel; K steps timed as one region, inputs come from host memory every step",
                    "per_call": {"value": e2e, "unit": "evals/s", "ms": e2e_tot / args.steps,
                                 "how": "depth=1: each call timed on its own (copy in, ONE kernel bk_forward_positions, copy out in stream order), L2 flushed between calls"}},
            "gpu_launches": launches_timed,
            "roofline": {"kernel": "bk_forward_tc_kernel", "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                         "frac": ach / tf_peak, "traffic": ncu_traffic("bk_forward_tc_kernel", B), "peak_source": peak_src,
                         "algorithmic_flop_per_launch": FLOP_VALID * B,
                         "achieved_dense_count": FLOP_DENSE * B / k_s / 1e12, "kernel_ms": 1e3 * k_s,
                         "one_launch_step": {"what": "the whole step as ONE launch (bk_forward_positions: planes computed on chip)",
                                             "ms": float(np.mean(t1_ms)), "evals_per_s": world * B / (1e-3 * float(np.mean(t1_ms)))},
                         "encoder": {"kernel": "bk_encode_kernel", "bound": "hbm", "ms": float(np.mean(e_ms)),
                                     "achieved": 4461 * B / (1e-3 * float(np.mean(e_ms))) / 1e9, "peak": hbm_peak, "unit": "GB/s"}},
        }
        if world == 1 and not args.no_cpu:
            rate, cores, ts = cpu_rate(args.cpu_sample, sd17, sdv, repeats=args.cpu_passes)      # ~10 s of CPU work
            if "selfplay" in extra:
                sp_rate, sp_cores, sp_dt = cpu_selfplay_rate(args.cpu_selfplay_games, sd17, sd19)
                extra["selfplay"]["cpu_baseline"] = {"value": sp_rate, "unit": "games/s", "cores": sp_cores, "kind": "port",
                                                     "sample": f"{args.cpu_selfplay_games} games x 72 moves in lock step ({sp_dt:.1f} s)"}
            if "simulate" in extra:
                sm_rate, sm_cores, sm_dt = cpu_simulate_rate(args.cpu_simulate_boards, sd17)
                extra["simulate"]["cpu_baseline"] = {"value": sm_rate, "unit": "playouts/s", "cores": sm_cores, "kind": "port",
                                                     "sample": f"{args.cpu_simulate_boards} boards played to the end in lock step ({sm_dt:.1f} s)"}
            if "mcts_genmove" in extra:
                g200, g1600 = cpu_genmove_reference(sd17, sdv, 200), cpu_genmove_reference(sd17, sdv, 1600)
                if g1600 is not None:
                    extra["mcts_genmove"]["cpu_baseline"] = dict(g1600, value=g1600["playouts_per_s"], unit="playouts/s")
                    if isinstance(extra.get("genmove_reference_callers"), dict):
                        extra["genmove_reference_callers"]["cpu_baseline"] = dict(g200, value=g200["playouts_per_s"], unit="playouts/s")
            line["cpu_baseline"] = {"value": rate, "unit": "evals/s", "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_sample} positions x {len(ts)} passes ({sum(ts):.1f} s); C feature oracle (OpenMP) + fp32 torch CPU forward, {cores} threads"}
        line.update(extra)
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=4096)
    ap.add_argument("--cpu-passes", type=int, default=8, help="passes of the CPU baseline over its sample (bounded: about 10 s)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-playouts", action="store_true", help="skip the self-play / simulate secondary measurements")
    ap.add_argument("--no-train", action="store_true", help="skip the REINFORCE step measurement")
    ap.add_argument("--selfplay-games", type=int, default=4096)
    ap.add_argument("--simulate-boards", type=int, default=65536)
    ap.add_argument("--cpu-selfplay-games", type=int, default=256)
    ap.add_argument("--cpu-simulate-boards", type=int, default=256)
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
