#!/bin/bash
# round 2, second session, batch 5: product-form time test in the latency-regime loop (variant prod) against the tree (PIPE + row gather)
V=$PWD/discretepomp.jl_b200/lib/variants
mkdir -p gpurun_out
DPOMP_LIB_PATH=$V/libdpomp_prod.so timeout 900 python -m pytest tests/test_gpu_pf.py tests/test_gpu_outer.py tests/test_gpu_callers.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2h_parity_prod.log
cat gpurun_out/r2h_parity_prod.log
for rep in 1 2; do
for v in base prod; do
  if [ "$v" = base ]; then unset DPOMP_LIB_PATH; else export DPOMP_LIB_PATH=$V/libdpomp_$v.so; fi
  echo "=== $v rep=$rep"
  python scripts/quick_bench.py pooley 200 1; python scripts/quick_bench.py pooley 200 64; python scripts/quick_bench.py pooley 200 256
  python scripts/quick_bench.py sir_c2 256 1; python scripts/quick_bench.py lotka_c4 256 16
done; done 2>&1 | tee gpurun_out/r2h_ab.log
