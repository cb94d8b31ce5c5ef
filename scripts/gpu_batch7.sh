#!/bin/bash
# round 2, second session, batch 4: row-wise gather of the resample kernel (variant gather = pipe + gather), bounds-checked build
V=$PWD/discretepomp.jl_b200/lib/variants
mkdir -p gpurun_out
DPOMP_LIB_PATH=$V/libdpomp_gather.so timeout 900 python -m pytest tests/test_gpu_pf.py tests/test_gpu_resample.py tests/test_gpu_outer.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2f_parity_gather.log
cat gpurun_out/r2f_parity_gather.log
timeout 900 python -m pytest tests/test_gpu_bounds.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2f_bounds.log
cat gpurun_out/r2f_bounds.log
for rep in 1 2; do
for v in base gather; do
  if [ "$v" = base ]; then unset DPOMP_LIB_PATH; else export DPOMP_LIB_PATH=$V/libdpomp_$v.so; fi
  echo "=== $v rep=$rep"
  python scripts/quick_bench.py sir_c2 1048576 1; python scripts/quick_bench.py seir_c3 65536 64
  if [ $rep = 1 ]; then python scripts/quick_bench.py lotka_c4 4096 1024; python scripts/quick_bench.py sir_dense 1048576 1; python scripts/quick_bench.py pooley 200 4000; fi
done; done 2>&1 | tee gpurun_out/r2f_ab.log
