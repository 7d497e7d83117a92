#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for the bench command; run under gpurun on ONE GPU.
#   tools/profile_gpu.sh <tag> [kernel-regex] [extra bench args]
# Writes gpurun_out/<tag>_launches.csv (every launch of one short chain with its device time) and
# gpurun_out/<tag>_full.ncu-rep (--set full capture of the first launches matching the regex).
TAG=${1:-prof}; REGEX=${2:-conv_tc}; shift 2
BENCH="python bench.py --precision bf16 --steps 1 --warmup 1 --T 2 --no-cpu-baseline $*"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${REGEX} -s 8 -c 6 -f -o gpurun_out/${TAG}_full $BENCH > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out/${TAG}_*
