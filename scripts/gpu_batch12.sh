#!/bin/bash
# round 2, second session: release-ordered ticket atomics (variant tkt2) against fence + atomicAdd (variant notkt); both with the
# deferred level 2 of the combine, no prefetch
V=$PWD/discretepomp.jl_b200/lib/variants
mkdir -p gpurun_out
DPOMP_LIB_PATH=$V/libdpomp_tkt2.so timeout 1200 python -m pytest tests/test_gpu_pf.py tests/test_gpu_resample.py tests/test_gpu_outer.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2m_parity_tkt2.log
cat gpurun_out/r2m_parity_tkt2.log
for rep in 1 2; do for v in notkt tkt2; do
  export DPOMP_LIB_PATH=$V/libdpomp_$v.so; echo "=== $v rep=$rep"
  python scripts/quick_bench.py sir_c2 1048576 1; python scripts/quick_bench.py seir_c3 65536 8; python scripts/quick_bench.py seir_c3 65536 16
  if [ $rep = 1 ]; then python scripts/quick_bench.py seir_c3 65536 64; python scripts/quick_bench.py lotka_c4 4096 1024; python scripts/quick_bench.py pooley 200 4000; python scripts/quick_bench.py pooley 200 1; fi
done; done 2>&1 | tee gpurun_out/r2m_ab.log
