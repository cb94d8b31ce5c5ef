import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpomp_b200 as dp
w = np.random.default_rng(0).random(8192)
for i in range(6):
    t0 = time.perf_counter(); idx = dp.rs_systematic(w, u=0.3); print(f"rs_systematic call {i}: {1e3*(time.perf_counter()-t0):.2f} ms")
model = dp.generate_model("LOTKA", [70, 70]); y = dp.get_observations("tests/golden/lotka_c4.csv")
hmm = dp.get_private_model(model, y)
pf = dp.ParticleFilter(dp.device_model(hmm), 4096, 8192, 1, seed=1)
pf2 = dp.ParticleFilter(dp.device_model(hmm), 4096, 8192, 1, seed=2)
for i in range(3):
    t0 = time.perf_counter(); idx = dp.rs_systematic(w, u=0.3); print(f"with big handles alive, call {i}: {1e3*(time.perf_counter()-t0):.2f} ms")
t0 = time.perf_counter(); pf.permute(idx); print(f"permute 8192 filters: {1e3*(time.perf_counter()-t0):.2f} ms")
t0 = time.perf_counter(); pf.copy_from(pf2, np.arange(1, 4097), np.arange(1, 4097)); print(f"copy 4096 filters: {1e3*(time.perf_counter()-t0):.2f} ms")
