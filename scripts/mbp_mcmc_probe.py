"""Time per lock-step MBP-MCMC step as a function of the number of chains (SIS / pooley.csv)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpomp_b200 as dp
model = dp.generate_model("SIS", [100, 1]); model.prior = dp.UniformProduct([0, 0], [0.01, 0.5])
y = dp.get_observations("tests/golden/pooley.csv")
hmm = dp.get_private_model(model, y)
acc = {}
def timed(cls, name):
    f = getattr(cls, name)
    def g(self, *a, **k):
        t = time.perf_counter(); r = f(self, *a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t; return r
    setattr(cls, name, g)
for nm in ("propose", "accept", "iterate", "set_stream_key"): timed(dp.MbpParticles, nm)
for n in (16, 1024, 4096, 16384):
    acc.clear()
    th0 = np.array([0.003, 0.1])[:, None] * np.random.default_rng(1).uniform(0.8, 1.2, size=(2, n))
    steps = 200
    t = time.time(); r = dp.run_mbp_mcmc(hmm, th0, steps, 100, False, seed=2, verbose=False); dt = time.time() - t
    print("  host-side split (s):", {k: round(v, 3) for k, v in acc.items()}, "total", round(dt, 3))
    pt = r.particles
    th = th0.copy()
    per_mode = []
    for mode in (1, 2):
        pt.set_mode(mode)
        pt.propose(th, th * 1.01, np.ones(n, dtype=bool), len(y))
        t = time.time()
        for _ in range(20): pt.propose(th, th * 1.01, np.ones(n, dtype=bool), len(y))
        per_mode.append((time.time() - t) / 20)
    print(f"chains {n}: {1e3*dt/steps:.2f} ms per MCMC step ({n*steps/dt:.3e} chain-steps/s); propose call alone: "
          f"thread per trajectory {1e3*per_mode[0]:.2f} ms, warp per trajectory {1e3*per_mode[1]:.2f} ms")
