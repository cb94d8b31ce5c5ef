#!/bin/bash
# One measurement pass on the GPU box (run through gpurun from the repo root):
#   scripts/measure_pass.sh TAG [skip-tests]
# Writes gpurun_out/{pytest_gpu,bench_ref,bench,launches}_TAG.* and gpurun_out/prof_TAG.ncu-rep;
# scripts/make_profiles.py TAG then turns them into the tracked summaries under profiles/.
TAG=${1:-r1}
mkdir -p gpurun_out
if [ "$2" != "skip-tests" ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_${TAG}.log
fi
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
python bench.py --impl reference > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 1
# profiler passes only after the plain command exited 0; numbers printed under ncu are never bench values
ncu --metrics gpu__time_duration.sum --clock-control none -c 440 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-smc2 --no-outer > gpurun_out/ncu1_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pf_ -s 210 -c 4 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-smc2 --no-outer > gpurun_out/ncu2_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_gpu_${TAG}.log 2>/dev/null; cat gpurun_out/smoke_${TAG}.log | tail -2; cat gpurun_out/bench_${TAG}.json | cut -c1-600
