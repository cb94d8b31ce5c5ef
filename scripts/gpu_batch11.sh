#!/bin/bash
# round 2, second session: measurement pass of the tree with the deferred level 2 + same-box A/B of the knob
bash scripts/measure_pass.sh r2d
for rep in 1 2; do for v in 0 1; do
  export DPOMP_DEFER_L2=$v; echo "=== defer_l2=$v rep=$rep"
  python scripts/quick_bench.py sir_c2 1048576 1; python scripts/quick_bench.py seir_c3 65536 8; python scripts/quick_bench.py seir_c3 65536 16
done; done 2>&1 | tee gpurun_out/r2l_ab.log
