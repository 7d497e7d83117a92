mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s3i_pytest.log 2>&1; tail -30 gpurun_out/s3i_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 1 --warmup 1 --T 10 --no-cpu-baseline --op-table gpurun_out/s3i_ops_lidc.txt > gpurun_out/s3i_lidc.json 2>&1
grep -i "attention" gpurun_out/s3i_ops_lidc.txt; tail -1 gpurun_out/s3i_ops_lidc.txt
timeout 300 python bench.py --workload cityscapes --steps 1 --warmup 1 --T 6 --no-cpu-baseline --op-table gpurun_out/s3i_ops_cs.txt > gpurun_out/s3i_cs.json 2>&1
grep -i "attention" gpurun_out/s3i_ops_cs.txt; tail -1 gpurun_out/s3i_ops_cs.txt
