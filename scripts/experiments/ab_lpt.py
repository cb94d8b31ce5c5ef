"""A/B of the heaviest-first launch order on heterogeneous LOTKA batches (the per-rank shape of SMC^2 C4 at 8 GPUs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dpomp_b200 as dp
model = dp.generate_model("LOTKA", [70, 70])
y = dp.get_observations("tests/golden/lotka_c4.csv")
dm = dp.device_model(dp.get_private_model(model, y))
theta = np.array([0.5, 0.0025, 0.3])
for nb in (1024, 2048):
    ths = theta[None, :] * np.random.default_rng(1).uniform(0.8, 1.25, (nb, 3))
    th = torch.tensor(ths, dtype=torch.float64, device="cuda"); out = torch.zeros(nb, dtype=torch.float64, device="cuda")
    for no_lpt in (True, False):
        if no_lpt: os.environ["DPOMP_NO_LPT"] = "1"
        pf = dp.ParticleFilter(dm, 4096, nb, 1, seed=1)
        os.environ.pop("DPOMP_NO_LPT", None)
        ms = []
        for _ in range(8):
            pf.set_stream_key(5); pf.loglik_device(th.data_ptr(), nb, out.data_ptr()); ms.append(pf.last_timing()[0])
        print(f"lotka heterogeneous nb={nb} x 4096, T=30: lpt={'off' if no_lpt else 'on '} {np.median(ms[2:]):.3f} ms (first call {ms[0]:.3f}); ll mean {out.mean().item():.4f}", flush=True)
        del pf
