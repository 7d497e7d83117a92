mkdir -p gpurun_out
timeout 60 python tools/tiny_chain.py bf16 128 8 > gpurun_out/s9b_tiny.log 2>&1 || { echo "tiny chain FAILED"; tail -5 gpurun_out/s9b_tiny.log; exit 1; }
timeout 300 python -m pytest tests -m gpu -q -s -k "head_fast" 2>&1 | grep -i "mismatch\|passed\|failed\|error" | head
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/s9b_pytest.log 2>&1; tail -3 gpurun_out/s9b_pytest.log | cut -c1-300
timeout 200 python bench.py --steps 2 --warmup 3 --T 40 --no-cpu-baseline --op-table gpurun_out/s9b_ops_lidc.txt > gpurun_out/s9b_lidc.json 2>&1
grep -n "head\|32->2 " gpurun_out/s9b_ops_lidc.txt | head; tail -1 gpurun_out/s9b_ops_lidc.txt
timeout 200 python bench.py --workload cityscapes --steps 2 --warmup 3 --T 20 --no-cpu-baseline --op-table gpurun_out/s9b_ops_cs.txt > gpurun_out/s9b_cs.json 2>&1
grep -n "head\|32->20 " gpurun_out/s9b_ops_cs.txt; tail -1 gpurun_out/s9b_ops_cs.txt
