#!/bin/bash
# Usage (under gpurun): scripts/ncu_capture.sh <tag> [bench args...]
# 1) plain run, 2) ncu launch list, 3) ncu --set full of the RL-iteration kernels.
# Output: gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep
set -u
TAG=$1; shift
B="python bench.py --steps 1 --warmup 1 --iterations 4 --no-cpu-baseline $*"
mkdir -p gpurun_out
$B > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_list_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:(row_(fast_)?kernel<\(int\)3|col_(fast_|sub_)?kernel<\(int\)1|col_(fast_|sub_)?kernel<\(int\)2|row_(fast_)?kernel<\(int\)4)' \
    -s 3 -c 5 -o gpurun_out/prof_$TAG $B > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out | tail -8
