#!/bin/bash
# one GPU-box visit: new tests, persistent-kernel A/B, no-evcount A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_cdriver.py tests/test_gpu_multi.py tests/test_reference_kats.py tests/test_gpu_bounds.py -m gpu -q -rs 2>&1 | tail -12
echo "=== mbp tests"; timeout 600 python -m pytest tests/test_gpu_mbp.py -m gpu -x -q 2>&1 | tail -6
echo "=== persistent tests"; timeout 600 python -m pytest tests/test_gpu_pf.py -m gpu -x -q -k "persistent" 2>&1 | tail -12
echo "=== persistent A/B"; timeout 600 python scripts/ab_persist.py 2>&1 | tee gpurun_out/r2e_persist.log
for v in "" noev; do
  if [ -n "$v" ]; then export DPOMP_LIB_PATH=$PWD/discretepomp.jl_b200/lib/variants/libdpomp_$v.so; fi
  echo "=== variant=$v"; python scripts/quick_bench.py sir_c2 1048576 1; python scripts/quick_bench.py seir_c3 65536 64; python scripts/quick_bench.py sir_c2 1048576 1
done 2>&1 | tee gpurun_out/r2d_noev.log
