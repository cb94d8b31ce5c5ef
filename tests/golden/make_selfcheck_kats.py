"""Writes tests/golden/ref_kats_selfcheck.json: vectors in the format of baseline/make_reference_kats.jl, but produced by
the ORACLE (oracle/), NOT by the reference.  They pin nothing about parity; they only keep the consumer
(tests/test_reference_kats.py) exercised until someone with a Julia installation generates the real file
tests/golden/ref_kats/ref_kats.json.  Run from the repository root:  python tests/golden/make_selfcheck_kats.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dpomp_b200 as dp  # noqa: E402
from oracle import oracle as orc  # noqa: E402

rng = np.random.default_rng(5)
out = {"julia_version": "none (oracle self-check)", "rs_systematic": [], "rs_stratified": [], "rs_multinomial": [], "rsp_systematic": [],
       "choose_event": [], "gom2": [], "rates": [], "moments": [], "pf_zero_rate": []}
wsets = [[1.0, 1.0, 1.0, 1.0], [0.0, 0.0, 1.0, 0.0], [3.0, 1.0], list(rng.random(64)), list(np.exp(-20 * rng.random(300)))]
for w in wsets:
    n = len(w)
    u = [float(rng.random())]
    out["rs_systematic"].append({"w": w, "u": u, "idx": orc.rs(1, w, u).tolist()})
    out["rsp_systematic"].append({"cw": np.cumsum(w).tolist(), "u": u, "idx": orc.rsp(1, np.cumsum(w), u).tolist()})
    u = rng.random(n).tolist()
    out["rs_stratified"].append({"w": w, "u": u, "idx": orc.rs(2, w, u).tolist()})
    out["rs_multinomial"].append({"w": w, "u": u, "idx": orc.rs(3, w, u).tolist()})
for _ in range(20):
    e = int(rng.integers(1, 7))
    cum = np.cumsum(rng.random(e) * (rng.random(e) > 0.25))
    if cum[-1] == 0:
        cum[:] = 1.0
    u = float(rng.random())
    out["choose_event"].append({"cum": cum.tolist(), "u": u, "event": orc.choose_event(cum, u)})
for sigma, seq in ((2.0, 2), (1.0, 2), (2.0, 3)):
    for yv, xv in (([0, 18, 0], [83, 18, 0]), ([0, 65, 7], [30, 71, 0])):
        d = sum(yv[seq - 1:seq]) - sum(xv[seq - 1:seq])
        out["gom2"].append({"sigma": sigma, "seq": seq, "y": yv, "x": xv,
                            "value": float(np.log(1 / (np.sqrt(2 * np.pi) * sigma)) - d * d / (2 * sigma * sigma))})
for name, ic in (("SIS", [100, 1]), ("SEIR", [100, 0, 1, 0]), ("LOTKA", [70, 70])):
    m = dp.generate_model(name, ic)
    e = m.m_transition.shape[0]
    for _ in range(3):
        theta = rng.random(max(e, 3)) * 0.1
        x = rng.integers(1, 300, len(ic))
        r = np.zeros(e)
        m.rate_function(r, theta, x)
        out["rates"].append({"model": name, "freq_dep": 0, "ic": ic, "theta": theta.tolist(), "x": x.tolist(), "rates": r.tolist(),
                             "trans": m.m_transition.reshape(-1).tolist()})
w = rng.random(50); th = rng.random((2, 50))
mu, cv = orc.compute_is_mu_covar(th, w)
out["moments"].append({"w": w.tolist(), "theta": th.T.reshape(-1).tolist(), "ess": orc.compute_ess(w), "mu": mu.tolist(), "cv": cv.reshape(-1).tolist()})
ys = [18, 65, 70, 66, 67]
ll = sum(np.log(1 / (np.sqrt(2 * np.pi) * 2.0)) - (v - 1) ** 2 / 8.0 for v in ys)
out["pf_zero_rate"].append({"model": "SIS", "ic": [100, 1], "np": 8, "loglik": float(ll)})
with open(os.path.join(ROOT, "tests", "golden", "ref_kats_selfcheck.json"), "w") as f:
    json.dump(out, f)
print("wrote ref_kats_selfcheck.json")
