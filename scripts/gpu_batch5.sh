#!/bin/bash
# round 2, second session, batch 2: parity of variant B, then a same-box A/B
#   noX = ticket tree + fused guess (best of batch 1);  A = noX + 32-bit clamps + AoS particle records
#   B = last-CTA combine (interleaved chains) + resampling uniform drawn by the combiner + everything of A;  C = B without AoS
V=$PWD/discretepomp.jl_b200/lib/variants
mkdir -p gpurun_out
DPOMP_LIB_PATH=$V/libdpomp_B.so timeout 900 python -m pytest tests/test_gpu_pf.py tests/test_gpu_resample.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2d_parity_B.log
DPOMP_LIB_PATH=$V/libdpomp_A.so timeout 900 python -m pytest tests/test_gpu_pf.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2d_parity_A.log
cat gpurun_out/r2d_parity_B.log gpurun_out/r2d_parity_A.log
for rep in 1 2; do
for v in base noX A B C; do
  if [ "$v" = base ]; then unset DPOMP_LIB_PATH; else export DPOMP_LIB_PATH=$V/libdpomp_$v.so; fi
  echo "=== $v rep=$rep"; python scripts/quick_bench.py sir_c2 1048576 1
  if [ $rep = 1 ] && [ $v != noX ]; then python scripts/quick_bench.py seir_c3 65536 64; python scripts/quick_bench.py lotka_c4 4096 1024; python scripts/quick_bench.py pooley 200 64; python scripts/quick_bench.py sir_dense 1048576 1; fi
done; done 2>&1 | tee gpurun_out/r2d_ab.log
DPOMP_LIB_PATH=$V/libdpomp_phase.so python scripts/phase_probe.py sir_c2 2>&1 | tee gpurun_out/r2d_phase.log
