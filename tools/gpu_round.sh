#!/bin/bash
# One GPU-box visit: parity tests, both benches, ncu launch list + one full capture. Outputs under gpurun_out/<tag>_*.
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_lidc.json 2> gpurun_out/${TAG}_bench_lidc.err; tail -c 3000 gpurun_out/${TAG}_bench_lidc.json
timeout 900 python bench.py --workload cityscapes --steps 1 --warmup 3 --cpu-budget 10 > gpurun_out/${TAG}_bench_cs.json 2> gpurun_out/${TAG}_bench_cs.err; tail -c 3000 gpurun_out/${TAG}_bench_cs.json; tail -3 gpurun_out/${TAG}_bench_cs.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2>&1; tail -c 1000 gpurun_out/${TAG}_bench_ref.json
timeout 900 tools/profile_gpu.sh ${TAG} conv_ws 20
