#!/bin/bash
# round 2, second session: two particles per lane + Philox drawn ahead in EVERY instantiation (variant ilp2, -DDPOMP_SIM_ILP=2)
# against the tree (one particle per lane in the throughput regime), on the event-heavy shapes
V=$PWD/discretepomp.jl_b200/lib/variants
for v in base ilp2; do
  if [ "$v" = base ]; then unset DPOMP_LIB_PATH; else export DPOMP_LIB_PATH=$V/libdpomp_$v.so; fi
  echo "=== $v"
  python scripts/quick_bench.py lotka_c4 4096 1024; python scripts/quick_bench.py sir_dense 1048576 1; python scripts/quick_bench.py seir_c3 65536 64
  python scripts/quick_bench.py sir_c2 1048576 1; python scripts/quick_bench.py pooley 200 4000
done 2>&1 | tee gpurun_out/r2q_ab.log
