"""A/B of the persistent cooperative kernel (dpomp_pf_set_persistent) against the per-observation launch chain."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dpomp_b200 as dp
cases = {"sir_c2": ("SIR", [100, 1, 0], [0.003, 0.1]), "sir_dense": ("SIR", [1000, 10, 0], [0.0003, 0.1]),
         "seir_c3": ("SEIR", [100, 0, 1, 0], [0.005, 0.2, 0.1]), "lotka_c4": ("LOTKA", [70, 70], [0.5, 0.0025, 0.3]),
         "pooley": ("SIS", [100, 1], [0.003, 0.1])}
runs = [("sir_c2", 1 << 20, 1), ("sir_c2", 1 << 19, 2), ("seir_c3", 65536, 16), ("lotka_c4", 4096, 256), ("sir_dense", 1 << 20, 1),
        ("pooley", 200, 1), ("pooley", 200, 64), ("pooley", 200, 1000)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for case, n, nb in runs:
    mname, ic, theta = cases[case]
    model = dp.generate_model(mname, ic)
    y = dp.get_observations(f"tests/golden/{case}.csv")
    dm = dp.device_model(dp.get_private_model(model, y))
    th = torch.tensor(np.tile(np.asarray(theta)[None, :], (nb, 1)), dtype=torch.float64, device="cuda")
    out = torch.zeros(nb, dtype=torch.float64, device="cuda")
    for mode in (0, 1):
        pf = dp.ParticleFilter(dm, n, nb, 1, seed=1)
        pf.set_persistent(mode)
        for _ in range(3): pf.loglik_device(th.data_ptr(), nb, out.data_ptr())
        ms = []
        for _ in range(12):
            flush.zero_(); torch.cuda.synchronize()
            pf.loglik_device(th.data_ptr(), nb, out.data_ptr()); ms.append(pf.last_timing()[0])
        print(f"{case} n={n} nb={nb} persistent={mode}: {np.median(ms):.3f} ms median ({min(ms):.3f} best) -> "
              f"{n*nb*len(y)/(np.median(ms)*1e-3):.3e} steps/s; launches {pf.last_timing()[1]}; ll={out.mean().item():.4f}", flush=True)
        del pf
