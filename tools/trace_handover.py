#!/usr/bin/env python
"""Where does a failing hand-over variant first go wrong?  Needs a measurement build with -DBK_TRACE (and, to see failures,
-DBK_HANDOVER=1): every epilogue thread records a hash of the accumulators it reads out of TMEM, per pass; a launch whose logits
differ from the first launch's is compared hash by hash: first pass, tiles (warp groups) and rows that deviate.
    BOKEGO_B200_SO=bokego_b200/libbokego_b200_tr1.so python tools/trace_handover.py [--iters 4000]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bokego_b200 import _lib, batched as bk  # noqa: E402

PASSES = 24


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=4000)
    ap.add_argument("--batch", type=int, default=741)
    ap.add_argument("--max-bad", type=int, default=12)
    a = ap.parse_args()
    g = os.path.join(ROOT, "tests", "golden")
    P, sd = dict(np.load(os.path.join(g, "positions.npz"))), dict(np.load(os.path.join(g, "weights_policy_17.npz")))
    dev = torch.device("cuda", 0)
    pol = bk.PackedNet(sd, dev)
    L = _lib.lib()
    so = C.CDLL(_lib.SO_PATH)
    so.bk_debug_trace.argtypes = [C.c_void_p]
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    trace = torch.zeros(n_sm, PASSES, 512, dtype=torch.int32, device=dev)
    assert so.bk_debug_trace(trace.data_ptr()) == 0
    B = a.batch
    idx = np.arange(B) % len(P["board"])
    pos = bk.Positions.from_numpy(P["board"][idx], P["ko"][idx], P["last"][idx], P["turn"][idx], dev)
    conv = bk.features_batch(pos, want=("conv",))["conv"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = ref_tr = None
    n_bad = n_odd = 0
    for it in range(a.iters):
        flush.zero_()
        trace.zero_()
        l, p, v = bk.policy_value_batch(conv, B, pol, pol)
        torch.cuda.synchronize()
        if ref is None:
            ref, ref_tr = l.clone(), trace.clone()
            continue
        if torch.equal(l, ref):
            if not torch.equal(trace, ref_tr) and n_odd < 5:
                n_odd += 1
                d = (trace != ref_tr)
                blocks = torch.nonzero(d.any(2).any(1)).flatten().tolist()
                print(f"iter {it}: logits equal but traces differ in {len(blocks)} blocks {blocks[:8]}")
                for b in blocks[:3]:
                    passes = torch.nonzero(d[b].any(1)).flatten().tolist()
                    thr = torch.nonzero(d[b, passes[0]]).flatten().tolist()
                    print(f"   block {b}: passes {passes}; pass {passes[0]}: {len(thr)} threads {thr[:10]}..{thr[-1]}")
            continue
        n_bad += 1
        rows = torch.nonzero((l != ref).any(1)).flatten().tolist()
        d = (trace != ref_tr)
        blocks = torch.nonzero(d.any(2).any(1)).flatten().tolist()
        print(f"iter {it}: {len(rows)} rows differ {rows[:16]}; blocks with deviating hashes {blocks}")
        for b in blocks:
            passes = torch.nonzero(d[b].any(1)).flatten().tolist()
            first = passes[0]
            thr = torch.nonzero(d[b, first]).flatten().tolist()
            groups = sorted({t // 128 for t in thr})
            print(f"   block {b} (pair {b // 2}, rank {b % 2}): deviating passes {passes}; first pass {first}: {len(thr)} threads, "
                  f"tiles {groups}, thread range {thr[0]}..{thr[-1]}; by tile " +
                  ", ".join(f"t{gq}:{sum(1 for t in thr if t // 128 == gq)}" for gq in groups))
            if len(passes) > 1:
                nxt = passes[1]
                thr2 = torch.nonzero(d[b, nxt]).flatten().tolist()
                print(f"      next deviating pass {nxt}: {len(thr2)} threads, tiles {sorted({t // 128 for t in thr2})}, range {thr2[0]}..{thr2[-1]}")
        if n_bad >= a.max_bad:
            break
    print(f"{it + 1} cold launches, {n_bad} wrong")


if __name__ == "__main__":
    main()
