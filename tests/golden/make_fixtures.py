"""Generates the synthetic observation fixtures of BASELINE.json configs C2-C5 with the CPU oracle's
restatement of gillespie_sim (src/hmm_sim.jl:86-102: observations at tmax/num_obs spacing, obs_id = 1, y.val = state).
Run from the repo root:  python tests/golden/make_fixtures.py   (deterministic: fixed Philox keys, recorded below)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dpomp_b200 as dp  # noqa: E402
from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CONFIGS = {
    # name: (model, initial condition, theta, tmax, num_obs, min_events)
    "sir_c2": ("SIR", [100, 1, 0], [0.003, 0.1], 100.0, 100, 50),
    "sir_dense": ("SIR", [1000, 10, 0], [0.0003, 0.1], 100.0, 100, 500),
    "seir_c3": ("SEIR", [100, 0, 1, 0], [0.005, 0.2, 0.1], 100.0, 100, 50),
    "lotka_c4": ("LOTKA", [70, 70], [0.5, 0.0025, 0.3], 30.0, 30, 50),
}


def main():
    for name, (mname, ic, theta, tmax, nobs, min_ev) in CONFIGS.items():
        model = dp.generate_model(mname, ic)
        times = np.arange(1, nobs + 1) * (tmax / nobs)
        y = [dp.Observation(float(t), 1, 1.0, np.zeros(len(ic), dtype=np.int64)) for t in times]
        cm = dp.compile_model(model, y)
        key = 20261018
        while True:
            states, ev = orc.gillespie_sim(cm.desc, theta, key=key)
            if ev >= min_ev:
                break
            key += 1
        path = os.path.join(HERE, f"{name}.csv")
        with open(path, "w") as f:
            f.write("time, " + ", ".join(f"val{i + 1}" for i in range(len(ic))) + "\n")
            for t, s in zip(times, states):
                f.write(f"{t}, " + ", ".join(str(int(v)) for v in s) + "\n")
        print(f"{name}: key={key} events={ev} final={states[-1].tolist()} -> {path}")


if __name__ == "__main__":
    main()
