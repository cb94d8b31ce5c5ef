"""Quick C2 timing (device-resident) with per-kernel CUDA-event breakdown; prints one line."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dpomp_b200 as dp
case = sys.argv[1] if len(sys.argv) > 1 else "sir_c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cases = {"sir_c2": ("SIR", [100, 1, 0], [0.003, 0.1]), "sir_dense": ("SIR", [1000, 10, 0], [0.0003, 0.1]),
         "seir_c3": ("SEIR", [100, 0, 1, 0], [0.005, 0.2, 0.1]), "lotka_c4": ("LOTKA", [70, 70], [0.5, 0.0025, 0.3]),
         "pooley": ("SIS", [100, 1], [0.003, 0.1])}
mname, ic, theta = cases[case]
model = dp.generate_model(mname, ic)
y = dp.get_observations(f"tests/golden/{case if case != 'pooley' else 'pooley'}.csv")
dm = dp.device_model(dp.get_private_model(model, y))
pf = dp.ParticleFilter(dm, n, nb, 1, seed=1)
if os.environ.get("FUSED") is not None: pf.set_fused(os.environ["FUSED"] == "1")
th = torch.tensor(np.tile(np.asarray(theta)[None, :], (nb, 1)), dtype=torch.float64, device="cuda")
out = torch.zeros(nb, dtype=torch.float64, device="cuda")
for _ in range(3): pf.loglik_device(th.data_ptr(), nb, out.data_ptr())
ms = []
for _ in range(10):
    pf.loglik_device(th.data_ptr(), nb, out.data_ptr()); ms.append(pf.last_timing()[0])
pf.set_kernel_timing(True)
pf.loglik_device(th.data_ptr(), nb, out.data_ptr())
(k0, k1), (n0, n1) = pf.last_kernel_timing()
best = min(ms)
print(f"{case} n={n} nb={nb} T={len(y)}: {np.median(ms):.3f} ms median ({best:.3f} best) -> {n*nb*len(y)/(np.median(ms)*1e-3):.3e} steps/s; "
      f"sim {1e3*k0/max(n0,1):.1f} us x{n0}, resample {1e3*k1/max(n1,1):.1f} us x{n1}; events {pf.last_event_count():.3e}; ll={out[0].item():.3f}")
