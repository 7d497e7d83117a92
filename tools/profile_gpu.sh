#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for the bench command; run under gpurun on ONE GPU.
#   tools/profile_gpu.sh <tag> [kernel-regex] [launch-skip] [extra bench args]
# Writes gpurun_out/<tag>_launches.csv (every launch of OUR kernels in one short chain with its device
# time; torch's setup kernels are filtered out by name) and gpurun_out/<tag>_full.ncu-rep (--set full
# capture of a few launches matching the regex, after skipping <launch-skip> matching launches).
TAG=${1:-prof}; REGEX=${2:-conv_ws}; SKIP=${3:-0}; shift 3
BENCH="python bench.py --precision bf16 --steps 1 --warmup 1 --T 2 --no-cpu-baseline --no-op-profile $*"
OURS='regex:conv_|attention_|head_kernel|time_table|onehot|nchw_to'
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${REGEX} -s ${SKIP} -c 4 -f -o gpurun_out/${TAG}_full $BENCH > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out/${TAG}_*
