mkdir -p gpurun_out
timeout 60 python tools/tiny_chain.py bf16 128 8 > gpurun_out/s6c_tiny.log 2>&1; rc=$?; echo "tiny rc=$rc"; tail -1 gpurun_out/s6c_tiny.log
if [ $rc -eq 0 ]; then
timeout 240 python -m pytest tests -m gpu -q > gpurun_out/s6c_pytest.log 2>&1; tail -3 gpurun_out/s6c_pytest.log | cut -c1-300
timeout 200 python bench.py --steps 2 --warmup 3 --T 40 --no-cpu-baseline --op-table gpurun_out/s6c_ops_lidc.txt > gpurun_out/s6c_lidc.json 2>&1
head -8 gpurun_out/s6c_ops_lidc.txt; tail -1 gpurun_out/s6c_ops_lidc.txt
timeout 200 python bench.py --workload cityscapes --steps 2 --warmup 3 --T 30 --no-cpu-baseline --op-table gpurun_out/s6c_ops_cs.txt > gpurun_out/s6c_cs.json 2>&1
tail -1 gpurun_out/s6c_ops_cs.txt
fi
