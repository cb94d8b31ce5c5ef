"""Small run of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpomp_b200 as dp
model = dp.generate_model("SIS", [100, 1]); y = dp.get_observations("tests/golden/pooley.csv")
hmm = dp.get_private_model(model, y); dm = dp.device_model(hmm)
theta = np.array([0.003, 0.1])
for n, nb in ((200, 3), (3000, 2)):
    for rs in (1, 2, 3):
        for prec in (dp._capi.SIM_F32, dp._capi.SIM_F64):
            pf = dp.ParticleFilter(dm, n, nb, rs, seed=1, sim_precision=prec)
            pf.set_record_ancestors(True)
            ll = pf.partial(np.tile(theta[:, None], (1, nb)), 1, 2)
            pf.permute(np.arange(nb, 0, -1))
print("pf ok", ll)
m2 = dp.generate_model("ROSSMAC", [50, 5, 60, 5]); y2 = [dp.Observation(float(t), 1, 1.0, [0, 5, 0, 0]) for t in (1.0, 2.0)]
pf = dp.ParticleFilter(dp.device_model(dp.get_private_model(m2, y2)), 1500, 2, 1, seed=2)
print("rossmac", pf.loglik(np.tile(np.array([[0.02], [0.05], [0.05], [0.1], [0.5], [0.5]]), (1, 2))))
print("hook", dp.rs_systematic(np.random.default_rng(0).random(5000), u=0.3)[:5], dp.rs_multinomial(np.ones(300), u=np.random.default_rng(1).random(300))[:3])
pt = dp.MbpParticles(dm, 100, 2048, seed=3)
th = np.tile(theta[:, None], (1, 100))
pt.iterate(th, 1, True); pt.iterate(th, 2, False)
ll = pt.propose(th, th * 1.1, np.ones(100, dtype=bool), 2); pt.accept([1, 5, 9]); pt.permute(np.sort(np.random.default_rng(2).integers(1, 101, 100)))
print("mbp ok", ll[:2])
