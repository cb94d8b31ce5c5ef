#!/bin/bash
# Multi-GPU pass on one box (run through `gpurun --gpus 8`): bench.py under torchrun at the given rank counts (default 8 2),
# each run's JSON line and stderr kept under gpurun_out/ (strong-scaling blocks smc2 / pmcmc_c3 / mbp_ibis_c5 inside).
TAG=${1:-r2}; shift
NS=${@:-8 2}
mkdir -p gpurun_out
for N in $NS; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err
  echo "N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_${N}gpu.json"))
    print("value", d["value"], "ms", d["ms_per_step"])
    for k in ("smc2", "pmcmc_c3", "mbp_ibis_c5"):
        x = d.get(k) or {}
        print(k, x.get("wall_s"), x.get("minus_log_evidence"), x.get("samples_sha16"), x.get("result_sha16"), x.get("rank0_phase_seconds"))
except Exception as e:
    print("no json:", e)
PY
  tail -3 gpurun_out/bench_${TAG}_${N}gpu.err
done
