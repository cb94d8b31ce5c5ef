"""First on-GPU smoke of the path: parity diagnostics against the oracle + a rough timing of config C2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpomp_b200 as dp
from oracle import oracle as orc

def model_and_obs(name, ic, csv):
    m = dp.generate_model(name, ic)
    y = dp.get_observations(csv)
    return m, y, dp.get_private_model(m, y)

m, y, hmm = model_and_obs("SIS", [100, 1], "tests/golden/pooley.csv")
dm = dp.device_model(hmm)
desc = dm.compiled.desc
theta = np.array([0.003, 0.1])
for n in (200, 1000, 5000):
    for rs in (1, 2, 3):
        pf = dp.ParticleFilter(dm, n, 1, rs, seed=7, sim_precision=dp._capi.SIM_F64)
        pf.set_record_ancestors(True)
        tile, items = pf.geometry()
        key = 123456789 + n
        pf.set_stream_key(key)
        ll = pf.partial(theta, 1, 3)[0]
        o_ll, o_lw, o_anc, o_ev, o_ovf, o_pop = orc.pf_partial(desc, theta, n, None, 1, 3, rs, key, 0, orc.MODE_DEVICE, tile, items)
        lw = pf.last_logw(); anc = pf.last_ancestors(); pop = pf.get_pop(1)
        print(f"n={n} rs={rs} tile={tile} ll gpu={ll:.12f} orc={o_ll:.12f} d={ll-o_ll:.2e} lw_maxdiff={np.max(np.abs(lw-o_lw)):.2e} "
              f"anc_mismatch={int(np.sum(anc!=o_anc))} pop_mismatch={int(np.sum(pop!=o_pop))} ev gpu={pf.last_event_count()} orc={o_ev}")
# f32 statistics
pf = dp.ParticleFilter(dm, 2000, 64, 1, seed=11)
lls = pf.loglik(np.tile(theta[:, None], (1, 64)))
print("f32 N=2000 x64: mean", lls.mean(), "sd", lls.std(), "(oracle anchor -15.69, sd 0.07-0.10)")
pf = dp.ParticleFilter(dm, 65536, 8, 1, seed=12)
lls = pf.loglik(np.tile(theta[:, None], (1, 8)))
print("f32 N=65536 x8: mean", lls.mean(), "sd", lls.std())

# C2 timing
m, y, hmm = model_and_obs("SIR", [100, 1, 0], "tests/golden/sir_c2.csv")
dm2 = dp.device_model(hmm)
th2 = np.array([0.003, 0.1])
for prec in (dp._capi.SIM_F32, dp._capi.SIM_F64):
    pf = dp.ParticleFilter(dm2, 1 << 20, 1, 1, seed=3, sim_precision=prec)
    for it in range(4):
        t0 = time.time(); ll = pf.loglik(th2)[0]; dt = time.time() - t0
        ms, nl = pf.last_timing()
        print(f"C2 prec={prec} ll={ll:.4f} wall={dt*1e3:.2f} ms dev={ms:.2f} ms launches={nl} events={pf.last_event_count()} "
              f"steps/s={(1<<20)*100/(ms*1e-3):.3e} ovf={pf.overflow_count()}")
o_ll, o_ev = orc.pf_loglik(dm2.compiled.desc, th2, 1 << 16, key=5, threads=orc.max_threads())
print("oracle C2 N=65536 ll", o_ll, "events", o_ev)
