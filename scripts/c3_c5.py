"""Full-size timings of BASELINE configs C3 (pMCMC, 64 chains x 65536 particles, SEIR) and C5 (MBP-IBIS, 16384 theta, SEIR,
stratified outer resampling) on ONE GPU (the configs name 8 GPUs; chains / theta-particles shard linearly)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpomp_b200 as dp
model = dp.generate_model("SEIR", [100, 0, 1, 0]); model.prior = dp.UniformProduct([0, 0, 0], [0.02, 1.0, 0.5])
y = dp.get_observations("tests/golden/seir_c3.csv")
hmm = dp.get_private_model(model, y)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
th0 = np.tile(np.array([[0.005], [0.2], [0.1]]), (1, 64)) * np.random.default_rng(1).uniform(0.8, 1.25, (3, 64))
dp.run_pmcmc(hmm, th0[:, :64], steps=3, adapt_period=2, p=65536, seed=1, verbose=False)
import gc; gc.collect()  # free the warm-up handle outside the timed region (cudaFree synchronises)
t0 = time.time(); res = dp.run_pmcmc(hmm, th0, steps=steps, adapt_period=steps // 2, p=65536, seed=2, verbose=False); dt = time.time() - t0
print(f"C3 pMCMC: 64 chains x 65536 particles x {len(y)} obs, {steps} MH steps in {dt:.2f} s -> {64*(steps-1)/dt:.1f} chain-steps/s, "
      f"{64*(steps-1)*65536*len(y)/dt:.3e} particle-obs steps/s; mean acceptance {res.accepted.mean()/(steps-1):.2f}")
t0 = time.time(); r = dp.run_mbp_ibis(hmm, model.prior.rand(16384, np.random.default_rng(3)), 0.5, 3, False, 1.002, seed=4, outer_rs=dp.rs_stratified, verbose=False); dt = time.time() - t0
print(f"C5 MBP-IBIS: 16384 theta, T={len(y)}, n_props 3, stratified: {dt:.2f} s -> {16384*len(y)/dt:.1f} theta-particle-obs/s; -ln p(y) {r.bme}; mu {r.mu}; AR {r.k_log[1]/max(r.k_log[0],1):.2f}")
