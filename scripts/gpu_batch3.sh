#!/bin/bash
# same-box A/B of the whole library across commits (C2 and the SEIR batch), three interleaved repetitions
V=$PWD/discretepomp.jl_b200/lib/variants
export DPOMP_LIB_LENIENT=1
for rep in 1 2 3; do
for v in c0start c1persist c2ilp c3outer HEAD; do
  if [ "$v" = HEAD ]; then unset DPOMP_LIB_PATH; else export DPOMP_LIB_PATH=$V/libdpomp_$v.so; fi
  echo "=== $v rep=$rep"; python scripts/quick_bench.py sir_c2 1048576 1; [ $rep = 1 ] && python scripts/quick_bench.py seir_c3 65536 64
done; done 2>&1 | tee gpurun_out/r2i_commits.log
