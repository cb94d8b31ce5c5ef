"""Weight collapse: an observation that only a few particles explain -> almost all offspring from one tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpomp_b200 as dp
model = dp.generate_model("SIR", [100, 1, 0], obs_error=0.25)
y = [dp.Observation(1.0, 1, 1.0, [0, 9, 0]), dp.Observation(2.0, 1, 1.0, [0, 9, 0])]
hmm = dp.get_private_model(model, y)
pf = dp.ParticleFilter(dp.device_model(hmm), 1 << 20, 1, 1, seed=3)
pf.set_record_ancestors(True); pf.set_kernel_timing(True)
for _ in range(3):
    pf.partial(np.array([0.003, 0.1]), 1, 1)
    (k0, k1), _n = pf.last_kernel_timing()
    anc = pf.last_ancestors()
    print(f"collapse: distinct ancestors {len(np.unique(anc))} of {1<<20}; sim {1e3*k0:.1f} us, resample {1e3*k1:.1f} us")
