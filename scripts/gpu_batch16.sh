#!/bin/bash
# round 2, second session: the two-per-lane loop on 1024-particle tiles, selected per call (tests + A/B of the knob)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pf.py -m gpu -q -x -k "two_per_lane or deferred_level2 or fused_step or f64_loop_is_bit_exact" 2>&1 | tail -5 > gpurun_out/r2s_tests.log
cat gpurun_out/r2s_tests.log
for v in 0 auto; do
  if [ $v = auto ]; then unset DPOMP_TWO_PER_LANE; else export DPOMP_TWO_PER_LANE=$v; fi
  echo "=== two_per_lane=$v"
  python scripts/quick_bench.py lotka_c4 4096 64; python scripts/quick_bench.py lotka_c4 4096 128; python scripts/quick_bench.py lotka_c4 4096 1024
  python scripts/quick_bench.py pooley 4096 32; python scripts/quick_bench.py sir_c2 1048576 1; python scripts/quick_bench.py seir_c3 65536 8
done 2>&1 | tee gpurun_out/r2s_ab.log
