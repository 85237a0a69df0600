"""The floating-point bar of the conv kernel, stated with a DISTRIBUTION instead of a fitted constant.

tests/golden/nets_big.npz (tests/golden/make_golden_nets_big.py, the unmodified reference on CPU fp32): 4,496 positions --
the survey's own 400-position set, 2,048 fresh and 2,048 carried-cache states of seeded random games -- with the logits of
policy_17 AND policy_19 and the stand-in ValueNet's values.

What bounds the error is the arithmetic the north-star prescribes (16-bit operands, fp32 accumulation), not the kernel: conv
weights and the activations handed from layer to layer are rounded to fp16 (2^-11 relative) at 7 + 6 places, and
oracle.nets.policy_logits_f16_operands -- a CPU emulation of exactly that arithmetic -- shows, against the fp32 reference on
this population, probabilities mean 2.1e-4 / p99 7.9e-4 / p99.9 1.2e-3 / max 3.4e-3 and logits p99.9 4.0e-2 / max 5.3e-2;
every layer and both operand kinds contribute evenly, so no cheap partial fix exists (hi+lo operands would double the MMAs).
On the survey's own 400 positions the emulation gives 7.4e-4 / 2.8e-2 -- SURVEY 8c's "1e-3 / 5e-2" was the maximum of that
small set, and the tail grows with the population.  The kernel is therefore held to

    probabilities   p99 <= 1e-3,  p99.9 <= 1.5e-3,  max <= 5e-3       (per position: worst of the 81 squares)
    logits          p99.9 <= 5e-2,  max <= 8e-2
    value           max <= 1e-3
    mean, p99, p99.9 <= 1.25 x (the single worst position: 1.5 x) the same statistic of the CPU emulation of the prescribed
                    arithmetic (+1e-4 / +2e-3 absolute).  Measured on the B200 (profiles/r02_precision_report.json):
                    policy_17 probabilities mean 2.10e-4 / p99 8.1e-4 / p99.9 1.26e-3 / max 3.9e-3 against the emulation's
                    2.06e-4 / 7.9e-4 / 1.22e-3 / 3.4e-3; the CUDA-core kernel over the same operands: 2.08e-4 / 7.8e-4 /
                    1.29e-3 / 3.5e-3; arg-max identical on 4,496 / 4,496 (policy_17) and 4,495 / 4,496 (policy_19: one
                    position whose fp32 top-2 margin is 2.8e-4, where the emulation flips as well)
    survey's set    probabilities max <= 1e-3 and logits max <= 5e-2, as SURVEY 8c states them
    arg-max         IDENTICAL to the fp32 reference, except at positions whose fp32 top-2 logit margin is smaller than twice
                    the logit error measured at that position (a tie at the precision of the operands); those are listed
                    and there may be at most 0.1 % of them (north-star: "identical argmax move", SURVEY 8c: ">= 99.9 %").
The measured distribution is printed (pytest -s) and written to gpurun_out/precision_report.json when that directory exists
(committed as profiles/r02_precision_report.json)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_LOGIT_P999, TOL_LOGIT_MAX = 5e-2, 8e-2
TOL_PROB_P99, TOL_PROB_P999, TOL_PROB_MAX, TOL_VALUE = 1e-3, 1.5e-3, 5e-3, 1e-3


@pytest.fixture(scope="module")
def big():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "nets_big.npz")))


def _evaluate(bk, dev, big, sd_p, sd_v, simt=False):
    """kernel outputs for all positions of the fixture: fresh positions in one encoder launch, carried-cache ones in another"""
    n = len(big["board"])
    logits = torch.empty(n, 81)
    probs = torch.empty(n, 81)
    value = torch.empty(n)
    pol = bk.PackedNet(sd_p, dev)
    val = bk.PackedNet(sd_v, dev) if sd_v is not None else None
    for carried in (False, True):
        idx = np.flatnonzero((big["set"] == 2) == carried)
        pos = bk.Positions.from_numpy(big["board"][idx], big["ko"][idx], big["last"][idx], big["turn"][idx], dev,
                                      big["libs_in"][idx] if carried else None)
        conv = bk.features_batch(pos, want=("conv", "libs"))["conv"]
        l, p, v = bk.policy_value_batch(conv, len(idx), pol, val, simt=simt)
        logits[idx], probs[idx] = l.cpu(), p.cpu()
        if v is not None:
            value[idx] = v.cpu()
    return logits, probs, value


def _dist(err):
    e = np.sort(np.asarray(err, np.float64))
    q = lambda f: float(e[min(len(e) - 1, int(np.ceil(f * len(e))) - 1)])
    return {"max": float(e[-1]), "p99.9": q(0.999), "p99": q(0.99), "p50": q(0.5), "mean": float(e.mean())}


def _emulated(big, sd):
    """error of the CPU emulation of the prescribed arithmetic against the fp32 reference, same positions"""
    from oracle import cpu as ocpu
    from oracle import nets as onets
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    feats = np.zeros((len(big["board"]), 27, 81), np.uint8)
    for carried in (False, True):
        m = (big["set"] == 2) == carried
        feats[m] = ocpu.features_batch(big["board"][m], big["ko"][m], big["last"][m], big["turn"][m],
                                       big["libs_in"][m] if carried else None)[0]
    x = onets.planes_to_float(feats)
    return torch.cat([onets.policy_logits_f16_operands(sd, x[i:i + 512]) for i in range(0, len(x), 512)])


def _check_net(name, logits, probs, want_logits, emu_logits, sets, report):
    want_l = torch.from_numpy(want_logits)
    want_p = torch.softmax(want_l, 1)
    el = (logits - want_l).abs().max(1).values.numpy()            # per position: worst square
    ep = (probs - want_p).abs().max(1).values.numpy()
    el_emu = (emu_logits - want_l).abs().max(1).values.numpy()
    ep_emu = (torch.softmax(emu_logits, 1) - want_p).abs().max(1).values.numpy()
    got_am, want_am = logits.argmax(1).numpy(), want_l.argmax(1).numpy()
    top2 = torch.topk(want_l, 2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1]).numpy()
    differ = np.flatnonzero(got_am != want_am)
    ties = [{"position": int(i), "fp32_top2_margin": float(margin[i]), "logit_err_here": float(el[i]),
             "reference_move": int(want_am[i]), "kernel_move": int(got_am[i])} for i in differ]
    r = {"positions": len(el), "logits": _dist(el), "probs": _dist(ep),
         "emulation_of_prescribed_arithmetic": {"logits": _dist(el_emu), "probs": _dist(ep_emu),
                                                "argmax_agreement": float((emu_logits.argmax(1).numpy() == want_am).mean())},
         "kernel_vs_emulation_logits": _dist((logits - emu_logits).abs().max(1).values.numpy()),
         "survey_400_set": {"logits_max": float(el[sets == 0].max()), "probs_max": float(ep[sets == 0].max())},
         "probs_by_set": {str(k): _dist(ep[sets == k]) for k in (0, 1, 2)},
         "argmax_agreement": float((got_am == want_am).mean()), "argmax_differences": ties,
         "smallest_fp32_top2_margin": float(margin.min()), "positions_with_margin_below_5e-2": int((margin < 5e-2).sum())}
    report[name] = r
    print(f"{name}: logits {r['logits']}\n{name}: probs  {r['probs']}\n{name}: emulation {r['emulation_of_prescribed_arithmetic']}\n"
          f"{name}: survey set {r['survey_400_set']}; argmax agreement {r['argmax_agreement']:.5f}, differences {ties}")
    bad = report.setdefault("failures", [])
    chk = lambda ok, what: None if ok else bad.append(f"{name}: {what}")
    chk(r["logits"]["p99.9"] <= TOL_LOGIT_P999 and r["logits"]["max"] <= TOL_LOGIT_MAX, "logit bar")
    chk(r["probs"]["p99"] <= TOL_PROB_P99 and r["probs"]["p99.9"] <= TOL_PROB_P999 and r["probs"]["max"] <= TOL_PROB_MAX, "prob bar")
    for k, f in (("mean", 1.25), ("p99", 1.25), ("p99.9", 1.25), ("max", 1.5)):    # as accurate as the prescribed arithmetic allows
        chk(r["probs"][k] <= f * r["emulation_of_prescribed_arithmetic"]["probs"][k] + 1e-4, f"probs {k} vs emulation")
        chk(r["logits"][k] <= f * r["emulation_of_prescribed_arithmetic"]["logits"][k] + 2e-3, f"logits {k} vs emulation")
    chk(r["survey_400_set"]["probs_max"] <= 1e-3 and r["survey_400_set"]["logits_max"] <= 5e-2, "survey set bar")
    for t in ties:        # a difference is admissible only where the reference itself has a tie at operand precision
        chk(t["fp32_top2_margin"] < 2 * t["logit_err_here"], f"argmax differs without a tie: {t}")
    chk(len(ties) <= max(1, len(el) // 1000), "too many argmax differences")
    return r


def test_error_distribution_policy_and_value(big, sd17, sd19, sd_value):
    from bokego_b200 import batched as bk
    dev = torch.device("cuda", 0)
    report = {"fixture": "tests/golden/nets_big.npz", "arithmetic": "fp16 operands, fp32 accumulation (tcgen05 kind::f16)",
              "sets": {"0": "survey's 400 fresh positions", "1": "2048 fresh states of random games", "2": "2048 carried-cache states"}}
    l17, p17, val = _evaluate(bk, dev, big, sd17, sd_value)
    _check_net("policy_17", l17, p17, big["logits17"], _emulated(big, sd17), big["set"], report)
    l19, p19, _ = _evaluate(bk, dev, big, sd19, None)
    _check_net("policy_19", l19, p19, big["logits19"], _emulated(big, sd19), big["set"], report)
    ev = (val - torch.from_numpy(big["value"])).abs().numpy()
    report["value"] = _dist(ev)
    print("value:", report["value"])
    if report["value"]["max"] > TOL_VALUE:
        report["failures"].append("value bar")
    want_p17 = torch.softmax(torch.from_numpy(big["logits17"]), 1)
    # the CUDA-core kernel over the same fp16 operands (fp32 FMA accumulation, no tensor core): the same tail => the error is
    # the operand rounding, not the tensor-core arithmetic
    ls, ps, _ = _evaluate(bk, dev, big, sd17, None, simt=True)
    eps = (ps - want_p17).abs().max(1).values.numpy()
    report["policy_17_probs_simt_kernel"] = _dist(eps)
    print("policy_17 probs, CUDA-core kernel:", report["policy_17_probs_simt_kernel"])
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "precision_report.json"), "w") as f:
            json.dump(report, f, indent=1)
    assert not report["failures"], report["failures"]
