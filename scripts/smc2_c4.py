"""SMC^2 on BASELINE config C4 (LOTKA [70,70], 8192 theta x 4096 state particles, prior U(0,(1,0.01,1))); single or
multi rank (torchrun).  Prints one line of timing + evidence."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dpomp_b200 as dp
outer_p = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
npf = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
comm = None
if world > 1:
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    comm = dp.Comm()
model = dp.generate_model("LOTKA", [70, 70])
model.prior = dp.UniformProduct([0, 0, 0], [1.0, 0.01, 1.0])
y = dp.get_observations("tests/golden/lotka_c4.csv")
dp.run_ibis_analysis(model, y[:12], np=max(outer_p // 8, 8 * world), npf=npf, seed=7, comm=comm, verbose=False)  # warm-up (CUDA / NCCL init)
import gc; gc.collect()
if world > 1:
    torch.distributed.barrier()
torch.cuda.synchronize()
t0 = time.time()
res = dp.run_ibis_analysis(model, y, np=outer_p, npf=npf, seed=1, comm=comm, verbose=False)
dt = time.time() - t0
if rank == 0:
    print(f"SMC2 C4 outer_p={outer_p} npf={npf} world={world}: {dt:.2f} s  theta-particle-obs/s={outer_p*len(y)/dt:.1f} "
          f"bme={res.bme} mu={res.mu} k_log={res.k_log}")
tm = res.timers
keys = sorted(tm)
vals = torch.tensor([tm[k] for k in keys] + [dt], dtype=torch.float64, device="cuda")
if world > 1:
    allv = [torch.zeros_like(vals) for _ in range(world)]
    torch.distributed.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    m = torch.stack(allv).cpu().numpy()
    for j, k in enumerate(keys + ["total"]):
        print(f"  {k:22s} min {m[:, j].min():7.3f}  mean {m[:, j].mean():7.3f}  max {m[:, j].max():7.3f} s over ranks")
if world > 1:
    torch.distributed.destroy_process_group()
