#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for the bench command; run under gpurun on ONE GPU.
#   tools/profile_gpu.sh <tag> <workload> <idx of dominant conv_tma launch> <idx of the 64->32 full-res conv_tma launch>
# (indices count conv_tma launches from the start of the process; tools/show_tc_config.py / DESIGN.md say how to get them)
# Writes gpurun_out/<tag>_launches.csv (per-launch metrics of OUR kernels over the first chains of the bench
# command: duration, DRAM bytes, tensor-pipe activity, instructions), gpurun_out/<tag>_ops.json (the op list, launch
# order) and --set full captures of the two conv launches and of the first attention launch.
TAG=${1:-prof}; WL=${2:-lidc}; SKIP_A=${3:-0}; SKIP_B=${4:-0}
PREC=${PREC:-exact}
BENCH="python bench.py --workload $WL --precision $PREC --headline-only --steps 1 --warmup 1 --T 2 --no-cpu-baseline --no-op-profile"
OURS='regex:conv_|attention_|head_kernel|time_table|encode_input'
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
mkdir -p gpurun_out
ncu --metrics $M --clock-control none -k "$OURS" -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH --dump-ops gpurun_out/${TAG}_ops.json > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tma -s ${SKIP_A} -c 1 -f -o gpurun_out/${TAG}_full_a $BENCH > gpurun_out/${TAG}_full_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tma -s ${SKIP_B} -c 1 -f -o gpurun_out/${TAG}_full_b $BENCH > gpurun_out/${TAG}_full_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 0 -c 1 -f -o gpurun_out/${TAG}_full_att $BENCH > gpurun_out/${TAG}_full_att.log 2>&1
ls -la gpurun_out/${TAG}_*
