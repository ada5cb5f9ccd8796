#!/bin/bash
# One GPU-box visit of round 2: bring-up check of the tensor-core correlator under a hard timeout (a hung mbarrier
# pipeline must not hold the box), parity margins, the GPU test suite, the bench.
# Usage (from the repo root, under gpurun):  bash tools/gpu_r2.sh <tag> [quick]
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 300 python tests/gpu_checks/gpu_tc_check.py > gpurun_out/tc_check_$TAG.log 2>&1; echo "tc_check rc=$?"; cat gpurun_out/tc_check_$TAG.log | tail -30
timeout 300 python tests/gpu_checks/gpu_margins.py > gpurun_out/margins_$TAG.log 2>&1; echo "margins rc=$?"; cat gpurun_out/margins_$TAG.log
if [ "$2" == "quick" ]; then exit 0; fi
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
