#!/bin/bash
# round 2, second session, batch 6: fused step kernel (mode 2) with the specialised resample phase, per-group flags and no
# arrival tickets when the launch is co-resident (variant fz), against the two-kernel chain
V=$PWD/discretepomp.jl_b200/lib/variants
mkdir -p gpurun_out
DPOMP_LIB_PATH=$V/libdpomp_fz.so timeout 900 python -m pytest tests/test_gpu_pf.py tests/test_gpu_outer.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2j_parity_fz.log
cat gpurun_out/r2j_parity_fz.log
for rep in 1 2; do
for v in base fz0 fz1; do
  unset DPOMP_LIB_PATH FUSED
  if [ "$v" = fz0 ]; then export DPOMP_LIB_PATH=$V/libdpomp_fz.so FUSED=0; fi
  if [ "$v" = fz1 ]; then export DPOMP_LIB_PATH=$V/libdpomp_fz.so FUSED=1; fi
  echo "=== $v rep=$rep"
  python scripts/quick_bench.py sir_c2 1048576 1; python scripts/quick_bench.py seir_c3 65536 8
  if [ $rep = 1 ]; then python scripts/quick_bench.py seir_c3 65536 64; python scripts/quick_bench.py lotka_c4 4096 1024; python scripts/quick_bench.py pooley 200 4000; python scripts/quick_bench.py sir_dense 1048576 1; fi
done; done 2>&1 | tee gpurun_out/r2j_ab.log
DPOMP_LIB_PATH=$V/libdpomp_fzphase.so FUSED=1 python scripts/phase_probe.py sir_c2 2>&1 | tee gpurun_out/r2j_phase_fused.log
