#!/bin/bash
# A/B timing of experiment builds of the correlator (make variant NAME=...): kernel time at the bench shape per library.
# Usage (repo root, under gpurun):  bash tools/gpu_ab.sh <batch> <name> [<name> ...]      ("base" = liblcs_b200.so)
B=$1; shift
for n in "$@"; do
  lib=$PWD/lte-cell-scanner_b200/liblcs_b200_$n.so
  [ "$n" == "base" ] && lib=$PWD/lte-cell-scanner_b200/liblcs_b200.so
  for rep in 1 2; do
    echo -n "$n: "; LCS_B200_LIB=$lib timeout 300 python tools/gpu_tc_prof.py $B 2>&1 | tail -1
  done
done
