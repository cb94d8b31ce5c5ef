#!/bin/bash
# round 2, second session: final measurement pass (deferred level 2, no prefetch) + the knob on the shapes around the one-wave limit
bash scripts/measure_pass.sh r2e
for v in 0 1; do
  export DPOMP_DEFER_L2=$v; echo "=== defer_l2=$v"
  python scripts/quick_bench.py sir_c2 1048576 1; python scripts/quick_bench.py seir_c3 65536 16; python scripts/quick_bench.py sir_c2 65536 18; python scripts/quick_bench.py lotka_c4 4096 256
done 2>&1 | tee gpurun_out/r2n_ab.log
