mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" > gpurun_out/s5e_att.log 2>&1; rc=$?; tail -5 gpurun_out/s5e_att.log | cut -c1-300
if [ $rc -eq 0 ]; then
timeout 240 python -m pytest tests -m gpu -q > gpurun_out/s5e_pytest.log 2>&1; tail -3 gpurun_out/s5e_pytest.log | cut -c1-250
timeout 200 python bench.py --steps 2 --warmup 3 --T 40 --no-cpu-baseline --op-table gpurun_out/s5e_ops_lidc.txt > gpurun_out/s5e_lidc.json 2>&1
grep attention gpurun_out/s5e_ops_lidc.txt; tail -1 gpurun_out/s5e_ops_lidc.txt
timeout 200 python bench.py --workload cityscapes --steps 2 --warmup 3 --T 30 --no-cpu-baseline --op-table gpurun_out/s5e_ops_cs.txt > gpurun_out/s5e_cs.json 2>&1
grep attention gpurun_out/s5e_ops_cs.txt; tail -1 gpurun_out/s5e_ops_cs.txt
true
true
fi
