"""Per-CTA phase timeline of one observation step of C2 (debug build with -DDPOMP_PHASE_TIMERS):
   python discretepomp.jl_b200/build.py --variant phase -DDPOMP_PHASE_TIMERS
   DPOMP_LIB_PATH=.../lib/variants/libdpomp_phase.so python scripts/phase_probe.py [case] [n] [nb]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dpomp_b200 as dp
case = sys.argv[1] if len(sys.argv) > 1 else "sir_c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cases = {"sir_c2": ("SIR", [100, 1, 0], [0.003, 0.1]), "seir_c3": ("SEIR", [100, 0, 1, 0], [0.005, 0.2, 0.1])}
mname, ic, theta = cases[case]
model = dp.generate_model(mname, ic); y = dp.get_observations(f"tests/golden/{case}.csv")
pf = dp.ParticleFilter(dp.device_model(dp.get_private_model(model, y)), n, nb, 1, seed=1)
fused = os.environ.get("FUSED") == "1"  # fused step kernel whenever the tiles of a filter are co-resident (mode 2)
if fused: pf.set_fused(2)
th = torch.tensor(np.tile(np.asarray(theta)[None, :], (nb, 1)), dtype=torch.float64, device="cuda")
out = torch.zeros(nb, dtype=torch.float64, device="cuda")
for _ in range(4): pf.loglik_device(th.data_ptr(), nb, out.data_ptr())
torch.cuda.synchronize()
lib = dp._capi.lib()
ncta = min(4096, nb * ((n + 1023) // 1024))
def read(fn):
    buf = np.zeros((2, 4096, 8), dtype=np.uint64)
    rc = getattr(lib, fn)(buf.ctypes.data_as(C.c_void_p)); assert rc == 0
    return buf.astype(np.int64)
sim_all = read("dpomp_debug_phases_sim")[0, :ncta, :8]
sim, nxt = sim_all[:, :7], sim_all[:, 7]
rs = read("dpomp_debug_phases_sim" if fused else "dpomp_debug_phases_rs")[1, :ncta, :5]
if fused: rs[:, 0] = rs[:, 1]; rs[:, 2] = rs[:, 1]  # fused: only "after the wait" (1), "counts + barrier" (3) and "gather done" (4) exist
t0 = sim[:, 0].min()
def stats(name, v):
    v = (v - t0) / 1e3
    print(f"  {name:34s} min {v.min():7.2f}  p10 {np.percentile(v,10):7.2f}  median {np.median(v):7.2f}  p90 {np.percentile(v,90):7.2f}  max {v.max():7.2f} us")
print(f"{case} n={n} nb={nb}: {ncta} CTAs; times relative to the first CTA start of the simulate kernel (observation 50)")
for i, nm in enumerate(["sim: CTA start", "sim: after grid dependency wait", "sim: tile staged", "sim: event loop done (warp 0)",
                        "sim: weights + table done", "sim: scan done, stores issued", "sim: tickets / combine done"]):
    stats(nm, sim[:, i])
for i, nm in enumerate(["resample: CTA start", "resample: after dependency wait", "resample: tile loaded", "resample: counts + barrier", "resample: gather done"]):
    stats(nm, rs[:, i])
stats("next simulate kernel: CTA start", nxt)
d = np.diff(sim, axis=1) / 1e3
print("  per-CTA phase durations (us), median / p90 / max:")
for i, nm in enumerate(["launch->wait", "staging", "event loop", "weights", "scan+stores", "tickets"]):
    print(f"    {nm:14s} {np.median(d[:, i]):6.2f} {np.percentile(d[:, i], 90):6.2f} {d[:, i].max():6.2f}")
dr = np.diff(rs, axis=1) / 1e3
for i, nm in enumerate(["launch->wait", "loads", "counts", "gather"]):
    print(f"    rs {nm:11s} {np.median(dr[:, i]):6.2f} {np.percentile(dr[:, i], 90):6.2f} {dr[:, i].max():6.2f}")
