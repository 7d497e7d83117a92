mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3e_pytest.log 2>&1; tail -15 gpurun_out/s3e_pytest.log
python bench.py --steps 1 --warmup 1 --T 10 --no-cpu-baseline --op-table gpurun_out/s3e_ops_lidc.txt > gpurun_out/s3e_lidc.json 2>&1
head -12 gpurun_out/s3e_ops_lidc.txt; tail -1 gpurun_out/s3e_ops_lidc.txt
