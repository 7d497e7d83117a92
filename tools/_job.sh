mkdir -p gpurun_out
timeout 60 python tools/tiny_chain.py bf16 128 8 > gpurun_out/s4m_tiny.log 2>&1; rc=$?; echo "tiny rc=$rc"; tail -1 gpurun_out/s4m_tiny.log
if [ $rc -eq 0 ]; then
timeout 240 python -m pytest tests -m gpu -q > gpurun_out/s4m_pytest.log 2>&1; tail -3 gpurun_out/s4m_pytest.log | cut -c1-250
timeout 120 python bench.py --steps 1 --warmup 1 --T 10 --no-cpu-baseline --op-table gpurun_out/s4m_ops_lidc.txt > gpurun_out/s4m_lidc.json 2>&1
grep -E "up|64->32 @128|skip @128" gpurun_out/s4m_ops_lidc.txt; tail -1 gpurun_out/s4m_ops_lidc.txt
timeout 120 python bench.py --workload cityscapes --steps 1 --warmup 1 --T 6 --no-cpu-baseline --op-table gpurun_out/s4m_ops_cs.txt > gpurun_out/s4m_cs.json 2>&1
tail -1 gpurun_out/s4m_ops_cs.txt
fi
