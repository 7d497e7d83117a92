#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for the DINO condition encoder; run under gpurun on ONE GPU.
# Writes gpurun_out/<tag>_launches.csv (per-launch metrics of every kernel of `bench.py --encoder-only`) and --set full
# captures of one vit_linear launch (384 -> 1536, GELU) and of one attention launch (T = 2049, 6 heads of 64).
TAG=${1:-enc}
BENCH="python bench.py --encoder-only --steps 1 --no-cpu-baseline"
OURS='regex:vit_|attention_'
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
mkdir -p gpurun_out
ncu --metrics $M --clock-control none -k "$OURS" -c 90 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
# launch order per block: ln, qkv, attention, proj, ln, fc1, fc2 -> the 3rd vit_linear launch is fc1 of block 0
ncu --set full --clock-control none --import-source on -k regex:vit_linear -s 2 -c 1 -f -o gpurun_out/${TAG}_full_linear $BENCH > gpurun_out/${TAG}_full_linear.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 0 -c 1 -f -o gpurun_out/${TAG}_full_att $BENCH > gpurun_out/${TAG}_full_att.log 2>&1
ls -la gpurun_out/${TAG}_*
