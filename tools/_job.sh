mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s4e_pytest.log 2>&1; tail -5 gpurun_out/s4e_pytest.log | cut -c1-250
timeout 300 python bench.py --steps 1 --warmup 1 --T 10 --no-cpu-baseline --op-table gpurun_out/s4e_ops_lidc.txt > gpurun_out/s4e_lidc.json 2>&1
head -14 gpurun_out/s4e_ops_lidc.txt; tail -1 gpurun_out/s4e_ops_lidc.txt
for NK in 128 64; do
CCDM_ATT_NK=$NK timeout 300 python bench.py --workload cityscapes --steps 1 --warmup 1 --T 6 --no-cpu-baseline --op-table gpurun_out/s4e_ops_cs_$NK.txt > gpurun_out/s4e_cs_$NK.json 2>&1
echo "NK=$NK"; grep attention gpurun_out/s4e_ops_cs_$NK.txt; tail -1 gpurun_out/s4e_ops_cs_$NK.txt
done
CCDM_ATT_NK=64 timeout 300 python bench.py --steps 1 --warmup 1 --T 10 --no-cpu-baseline --op-table gpurun_out/s4e_ops_lidc_nk64.txt > /dev/null 2>&1; grep attention gpurun_out/s4e_ops_lidc_nk64.txt
