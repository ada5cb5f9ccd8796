#!/bin/bash
# One GPU-box visit: tests, margins, bench (both kernels), launch list, full ncu captures of the top kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
python tests/gpu_checks/gpu_margins.py > gpurun_out/margins_$TAG.log 2>&1; cat gpurun_out/margins_$TAG.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
python bench.py --steps 10 --warmup 3 --kernel fp32 --no-cpu-baseline > gpurun_out/bench_fp32_$TAG.json 2>> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_fp32_$TAG.json
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; cat gpurun_out/bench_ref_$TAG.json
python bench.py --workload tracker --batch 256 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tracker_$TAG.json 2>> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_tracker_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xcorr_fold_tc -s 3 -c 2 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_bench_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xcorr_fold_fp32 -s 3 -c 1 -f -o gpurun_out/prof_fp32_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --kernel fp32 > gpurun_out/ncu_full_bench_fp32_$TAG.log 2>&1
ls -la gpurun_out | tail -20
