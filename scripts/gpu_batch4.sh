#!/bin/bash
# round 2, second session: parity of the new combine / fused guess, then a same-box A/B of the variant libraries
#   all  = last-CTA combine + fused systematic guess + resample carve-out; noX / noG / noC = all minus one of them
V=$PWD/discretepomp.jl_b200/lib/variants
mkdir -p gpurun_out
DPOMP_LIB_PATH=$V/libdpomp_lastcta.so timeout 900 python -m pytest tests/test_gpu_pf.py tests/test_gpu_resample.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2c_parity.log
cat gpurun_out/r2c_parity.log
for rep in 1 2; do
for v in base lastcta noX noG noC; do
  if [ "$v" = base ]; then unset DPOMP_LIB_PATH; else export DPOMP_LIB_PATH=$V/libdpomp_$v.so; fi
  echo "=== $v rep=$rep"; python scripts/quick_bench.py sir_c2 1048576 1
  if [ $rep = 1 ]; then python scripts/quick_bench.py seir_c3 65536 64; fi
  if [ $rep = 1 ] && { [ $v = base ] || [ $v = lastcta ]; }; then python scripts/quick_bench.py lotka_c4 4096 1024; python scripts/quick_bench.py pooley 200 64; fi
done; done 2>&1 | tee gpurun_out/r2c_ab.log
DPOMP_LIB_PATH=$V/libdpomp_phase.so python scripts/phase_probe.py sir_c2 2>&1 | tee gpurun_out/r2c_phase.log
