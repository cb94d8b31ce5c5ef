#!/bin/bash
# round 2, second session, batch 3: the strengthened f32 per-trajectory test, and the Philox-ahead latency-regime loop (variant pipe)
V=$PWD/discretepomp.jl_b200/lib/variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pf.py -m gpu -q -s -k "f32_loop_trajectories" 2>&1 | tail -15 > gpurun_out/r2e_f32traj.log
cat gpurun_out/r2e_f32traj.log
DPOMP_LIB_PATH=$V/libdpomp_pipe.so timeout 900 python -m pytest tests/test_gpu_pf.py tests/test_gpu_outer.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2e_parity_pipe.log
cat gpurun_out/r2e_parity_pipe.log
for rep in 1 2; do
for v in base pipe; do
  if [ "$v" = base ]; then unset DPOMP_LIB_PATH; else export DPOMP_LIB_PATH=$V/libdpomp_$v.so; fi
  echo "=== $v rep=$rep"
  python scripts/quick_bench.py pooley 200 1; python scripts/quick_bench.py pooley 200 64; python scripts/quick_bench.py pooley 200 256
  python scripts/quick_bench.py sir_c2 256 1; python scripts/quick_bench.py lotka_c4 256 16
  if [ $rep = 1 ]; then python scripts/quick_bench.py pooley 200 4000; python scripts/quick_bench.py sir_c2 1048576 1; fi
done; done 2>&1 | tee gpurun_out/r2e_ab.log
