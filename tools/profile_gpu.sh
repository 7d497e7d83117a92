#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for the bench command; run under gpurun on ONE GPU.
#   tools/profile_gpu.sh <tag> <workload> <index of the dominant op among conv_tma launches of a step>
# Writes gpurun_out/<tag>_launches.csv (per-launch metrics of OUR kernels over the first chains of the bench
# command: duration, DRAM bytes, tensor-pipe activity), gpurun_out/<tag>_ops.json (the op list, launch order) and
# gpurun_out/<tag>_full.ncu-rep (--set full capture of the dominant kernel).
TAG=${1:-prof}; WL=${2:-lidc}; SKIP=${3:-0}
BENCH="python bench.py --workload $WL --precision bf16 --steps 1 --warmup 1 --T 2 --no-cpu-baseline --no-op-profile"
OURS='regex:conv_|attention_|head_kernel|time_table|encode_input'
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
mkdir -p gpurun_out
ncu --metrics $M --clock-control none -k "$OURS" -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH --dump-ops gpurun_out/${TAG}_ops.json > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tma -s ${SKIP} -c 1 -f -o gpurun_out/${TAG}_full $BENCH > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out/${TAG}_*
