mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q > gpurun_out/s6b_pytest.log 2>&1; tail -3 gpurun_out/s6b_pytest.log | cut -c1-300
for M in 1 2; do
CCDM_SILU_MODE=$M timeout 200 python bench.py --steps 2 --warmup 3 --T 40 --no-cpu-baseline --op-table gpurun_out/s6b_ops_lidc_$M.txt > gpurun_out/s6b_lidc_$M.json 2>&1
echo "mode $M"; head -6 gpurun_out/s6b_ops_lidc_$M.txt; tail -1 gpurun_out/s6b_ops_lidc_$M.txt
done
