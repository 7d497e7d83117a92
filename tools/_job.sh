mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s3k_pytest.log 2>&1; tail -5 gpurun_out/s3k_pytest.log | cut -c1-250
timeout 300 python bench.py --steps 1 --warmup 1 --T 10 --no-cpu-baseline --op-table gpurun_out/s3k_ops_lidc.txt > gpurun_out/s3k_lidc.json 2>&1
head -40 gpurun_out/s3k_ops_lidc.txt; tail -1 gpurun_out/s3k_ops_lidc.txt; tail -3 gpurun_out/s3k_lidc.json | cut -c1-300
