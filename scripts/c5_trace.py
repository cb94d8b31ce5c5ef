"""MBP-IBIS C5 under torchrun with the migration trace (DPOMP_TRACE_MIGRATE=1): per-segment host times of
dpomp_mbp_resample_migrate on rank 0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import dpomp_b200 as dp
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
if rank != 0:
    os.environ.pop("DPOMP_TRACE_MIGRATE", None)
model = dp.generate_model("SEIR", [100, 0, 1, 0]); model.prior = dp.UniformProduct([0, 0, 0], [0.02, 1.0, 0.5])
y = dp.get_observations("tests/golden/seir_c3.csv")
hmm = dp.get_private_model(model, y)
comm = dp.Comm() if world > 1 else None
th0 = model.prior.rand(16384, np.random.default_rng(3))
for rep in range(2):
    t0 = time.perf_counter()
    r = dp.run_mbp_ibis(hmm, th0, 0.5, 3, False, 1.002, seed=4, comm=comm, outer_rs=dp.rs_stratified, verbose=False)
    if rank == 0:
        print(f"rep {rep}: {time.perf_counter() - t0:.3f} s  bme {r.bme}  timers {({k: round(v, 4) for k, v in r.timers.items()})}", file=sys.stderr)
if world > 1:
    dist.destroy_process_group()
