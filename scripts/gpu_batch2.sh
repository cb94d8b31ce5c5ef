#!/bin/bash
V=$PWD/discretepomp.jl_b200/lib/variants
for rep in 1 2; do
for v in "" evred; do
  if [ -n "$v" ]; then export DPOMP_LIB_PATH=$V/libdpomp_$v.so; else unset DPOMP_LIB_PATH; fi
  echo "=== variant=${v:-main} rep=$rep"; python scripts/quick_bench.py sir_c2 1048576 1; python scripts/quick_bench.py seir_c3 65536 64; python scripts/quick_bench.py lotka_c4 4096 1024
done; done 2>&1 | tee gpurun_out/r2f_ev.log
for v in "" ilp2; do
  if [ -n "$v" ]; then export DPOMP_LIB_PATH=$V/libdpomp_$v.so; else unset DPOMP_LIB_PATH; fi
  echo "=== small filters variant=${v:-main}"; python scripts/quick_bench.py pooley 200 1; python scripts/quick_bench.py pooley 200 64; python scripts/quick_bench.py pooley 200 4000; python scripts/quick_bench.py sir_c2 256 1
done 2>&1 | tee gpurun_out/r2f_ilp.log
