#!/bin/bash
# round 2, second session, batch 7: level 2 of the combine deferred to the resample kernel (variant dl2; DPOMP_DEFER_L2=0 switches it off)
V=$PWD/discretepomp.jl_b200/lib/variants
mkdir -p gpurun_out
DPOMP_LIB_PATH=$V/libdpomp_dl2.so timeout 1200 python -m pytest tests/test_gpu_pf.py tests/test_gpu_resample.py tests/test_gpu_outer.py tests/test_gpu_callers.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2k_parity_dl2.log
cat gpurun_out/r2k_parity_dl2.log
export DPOMP_LIB_PATH=$V/libdpomp_dl2.so
for rep in 1 2; do
for v in 0 1; do
  export DPOMP_DEFER_L2=$v
  echo "=== defer_l2=$v rep=$rep"
  python scripts/quick_bench.py sir_c2 1048576 1; python scripts/quick_bench.py seir_c3 65536 8
  if [ $rep = 1 ]; then python scripts/quick_bench.py seir_c3 65536 64; python scripts/quick_bench.py lotka_c4 4096 1024; python scripts/quick_bench.py sir_dense 1048576 1; python scripts/quick_bench.py sir_c2 100000 16; fi
done; done 2>&1 | tee gpurun_out/r2k_ab.log
unset DPOMP_DEFER_L2
DPOMP_LIB_PATH=$V/libdpomp_dl2phase.so python scripts/phase_probe.py sir_c2 2>&1 | tee gpurun_out/r2k_phase.log
