"""A/B of the offspring placement (dpomp_pf_set_scatter) and of library variants on the timed configurations.
usage: python scripts/ab_scatter.py [case n nb]..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dpomp_b200 as dp
cases = {"sir_c2": ("SIR", [100, 1, 0], [0.003, 0.1]), "sir_dense": ("SIR", [1000, 10, 0], [0.0003, 0.1]),
         "seir_c3": ("SEIR", [100, 0, 1, 0], [0.005, 0.2, 0.1]), "lotka_c4": ("LOTKA", [70, 70], [0.5, 0.0025, 0.3])}
runs = [("sir_c2", 1 << 20, 1), ("seir_c3", 65536, 64), ("lotka_c4", 4096, 1024), ("sir_dense", 1 << 20, 1)]
if len(sys.argv) > 3:
    runs = [(sys.argv[i], int(sys.argv[i + 1]), int(sys.argv[i + 2])) for i in range(1, len(sys.argv) - 2, 3)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for case, n, nb in runs:
    mname, ic, theta = cases[case]
    model = dp.generate_model(mname, ic)
    y = dp.get_observations(f"tests/golden/{case}.csv")
    dm = dp.device_model(dp.get_private_model(model, y))
    rng = np.random.default_rng(1)
    ths = np.tile(np.asarray(theta)[None, :], (nb, 1)) * (rng.uniform(0.8, 1.25, (nb, len(theta))) if os.environ.get("HETERO") else 1.0)
    th = torch.tensor(ths, dtype=torch.float64, device="cuda")
    out = torch.zeros(nb, dtype=torch.float64, device="cuda")
    for mode in (0, 1):
        pf = dp.ParticleFilter(dm, n, nb, 1, seed=1)
        pf.set_scatter(mode)
        for _ in range(3): pf.loglik_device(th.data_ptr(), nb, out.data_ptr())
        ms = []
        for _ in range(12):
            flush.zero_(); torch.cuda.synchronize()
            pf.loglik_device(th.data_ptr(), nb, out.data_ptr()); ms.append(pf.last_timing()[0])
        pf.set_kernel_timing(True)
        pf.loglik_device(th.data_ptr(), nb, out.data_ptr())
        (k0, k1), (n0, n1) = pf.last_kernel_timing()
        print(f"{case} n={n} nb={nb} scatter={mode}: {np.median(ms):.3f} ms median ({min(ms):.3f} best) -> {n*nb*len(y)/(np.median(ms)*1e-3):.3e} steps/s; "
              f"sim {1e3*k0/max(n0,1):.1f} us x{n0}, resample {1e3*k1/max(n1,1):.1f} us x{n1}; ll={out.mean().item():.3f}", flush=True)
        del pf
