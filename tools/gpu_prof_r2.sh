#!/bin/bash
# Profiling visit: per-stage wait counters of the tcgen05 correlator, launch list of the bench step, full ncu captures.
# Usage (repo root, under gpurun):  bash tools/gpu_prof_r2.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
[ -f lte-cell-scanner_b200/liblcs_b200_prof.so ] || make -C lte-cell-scanner_b200 prof > /dev/null 2>&1   # instrumented build (stage counters)
LCS_B200_LIB=$PWD/lte-cell-scanner_b200/liblcs_b200_prof.so LCS_TC_PROF=1 timeout 300 python tools/gpu_tc_prof.py 384 > gpurun_out/tc_stage_waits_$TAG.txt 2>&1
cat gpurun_out/tc_stage_waits_$TAG.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_launch_bench_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:xcorr_fold_tc -s 3 -c 2 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_full_bench_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"epilogue4|sp_partial" -s 6 -c 2 -f -o gpurun_out/prof_epi_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_full_epi_$TAG.log 2>&1
ls -la gpurun_out | tail -8
