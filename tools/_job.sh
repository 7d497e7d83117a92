mkdir -p gpurun_out
for M in 3 1; do
CCDM_SILU_MODE=$M timeout 240 python -m pytest tests -m gpu -q -k "bf16 or tc_conv" > gpurun_out/s6e_pytest_$M.log 2>&1; tail -2 gpurun_out/s6e_pytest_$M.log | cut -c1-300
cp gpurun_out/parity_report.json gpurun_out/s6e_parity_$M.json
CCDM_SILU_MODE=$M timeout 200 python bench.py --steps 2 --warmup 3 --T 40 --no-cpu-baseline --op-table gpurun_out/s6e_ops_lidc_$M.txt > gpurun_out/s6e_lidc_$M.json 2>&1
echo "mode $M"; head -7 gpurun_out/s6e_ops_lidc_$M.txt; tail -1 gpurun_out/s6e_ops_lidc_$M.txt
done
