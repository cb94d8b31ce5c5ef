#!/bin/bash
# round 2, second session: is the two-particles-per-lane + Philox-ahead loop worth instantiating for 1024-particle tiles when the
# launch under-fills the device (the per-rank shape of C3 on 8 GPUs: 512 CTAs; a single 2^18 / 2^19-particle filter)?
V=$PWD/discretepomp.jl_b200/lib/variants
for v in base ilp2; do
  if [ "$v" = base ]; then unset DPOMP_LIB_PATH; else export DPOMP_LIB_PATH=$V/libdpomp_$v.so; fi
  echo "=== $v"
  python scripts/quick_bench.py seir_c3 65536 8; python scripts/quick_bench.py seir_c3 65536 4; python scripts/quick_bench.py seir_c3 65536 2
  python scripts/quick_bench.py sir_c2 262144 1; python scripts/quick_bench.py sir_c2 524288 1; python scripts/quick_bench.py lotka_c4 4096 64
done 2>&1 | tee gpurun_out/r2r_ab.log
