"""Fixed overhead of the two PF kernels: theta = 0 means no events at all (every particle finishes at its first rate
evaluation), so the simulate kernel's time is staging + work queue + weight/scan epilogue only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dpomp_b200 as dp
model = dp.generate_model("SIR", [100, 1, 0]); y = dp.get_observations("tests/golden/sir_c2.csv")
dm = dp.device_model(dp.get_private_model(model, y))
n = 1 << 20
for scale in (0.0, 0.25, 1.0, 4.0):
    pf = dp.ParticleFilter(dm, n, 1, 1, seed=1)
    th = torch.tensor([[0.003 * scale, 0.1 * scale]], dtype=torch.float64, device="cuda")
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    for _ in range(3): pf.loglik_device(th.data_ptr(), 1, out.data_ptr())
    ms = []
    for _ in range(5):
        pf.loglik_device(th.data_ptr(), 1, out.data_ptr()); ms.append(pf.last_timing()[0])
    ev = pf.last_event_count()
    pf.set_kernel_timing(True); pf.loglik_device(th.data_ptr(), 1, out.data_ptr())
    (k0, k1), (n0, n1) = pf.last_kernel_timing()
    print(f"theta x{scale}: {np.median(ms):.3f} ms, events/particle-step {ev/(n*100):.2f}, sim {1e3*k0/n0:.1f} us, resample {1e3*k1/n1:.1f} us")
