#!/bin/bash
# One GPU-box visit: parity tests, both benches (default flags = what the driver runs), the reference arm, and the
# ncu recipe (launch list + one --set full capture of the dominant kernel).  Outputs under gpurun_out/<tag>_*.
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --op-table gpurun_out/${TAG}_ops_lidc.txt > gpurun_out/${TAG}_bench_lidc.json 2> gpurun_out/${TAG}_bench_lidc.err; tail -c 2500 gpurun_out/${TAG}_bench_lidc.json; tail -3 gpurun_out/${TAG}_bench_lidc.err
timeout 900 python bench.py --workload cityscapes --steps 2 --warmup 3 --cpu-budget 10 --op-table gpurun_out/${TAG}_ops_cs.txt > gpurun_out/${TAG}_bench_cs.json 2> gpurun_out/${TAG}_bench_cs.err; tail -c 2500 gpurun_out/${TAG}_bench_cs.json; tail -3 gpurun_out/${TAG}_bench_cs.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2>&1; tail -c 700 gpurun_out/${TAG}_bench_ref.json
timeout 900 tools/profile_gpu.sh ${TAG}_lidc lidc 2 79
timeout 900 tools/profile_gpu.sh ${TAG}_cs cityscapes 4 103
timeout 300 python bench.py --precision fp32 --steps 1 --warmup 3 --T 10 --no-cpu-baseline --headline-only > gpurun_out/${TAG}_bench_lidc_fp32_T10.json 2>&1; tail -c 600 gpurun_out/${TAG}_bench_lidc_fp32_T10.json
