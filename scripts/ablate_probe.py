import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dpomp_b200 as dp
model = dp.generate_model("SIR", [100, 1, 0]); y = dp.get_observations("tests/golden/sir_c2.csv")
dm = dp.device_model(dp.get_private_model(model, y))
n = 1 << 20
pf = dp.ParticleFilter(dm, n, 1, 1, seed=1)
th = torch.tensor([[0.0, 0.0]], dtype=torch.float64, device="cuda"); out = torch.zeros(1, dtype=torch.float64, device="cuda")
for _ in range(3): pf.loglik_device(th.data_ptr(), 1, out.data_ptr())
pf.set_kernel_timing(True); pf.loglik_device(th.data_ptr(), 1, out.data_ptr())
(k0, k1), (n0, n1) = pf.last_kernel_timing()
print(f"ablate={os.environ.get('ABL','0')}: sim {1e3*k0/n0:.1f} us, resample {1e3*k1/n1:.1f} us")
