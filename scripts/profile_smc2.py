import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpomp_b200 as dp
model = dp.generate_model("LOTKA", [70, 70]); model.prior = dp.UniformProduct([0, 0, 0], [1.0, 0.01, 1.0])
y = dp.get_observations("tests/golden/lotka_c4.csv")
dp.run_ibis_analysis(model, y[:3], np=256, npf=4096, seed=3, verbose=False)
pr = cProfile.Profile(); t0 = time.time(); pr.enable()
res = dp.run_ibis_analysis(model, y, np=8192, npf=4096, seed=1, verbose=False)
pr.disable(); print("wall", time.time() - t0, res.bme)
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
