mkdir -p gpurun_out
PREV=$PWD/ccdm-stochastic-segmentation_b200/ccdm_b200/libccdm_b200_prev.so
timeout 60 python tools/tiny_chain.py bf16 128 8 > gpurun_out/s12a_tiny.log 2>&1 || { echo "tiny chain FAILED"; tail -5 gpurun_out/s12a_tiny.log; exit 1; }
timeout 200 python -m pytest tests -m gpu -q -x -k "attention" > gpurun_out/s12a_pytest.log 2>&1; tail -1 gpurun_out/s12a_pytest.log | cut -c1-200
for V in prev new; do
case $V in
prev) export CCDM_B200_LIB=$PREV;;
new) unset CCDM_B200_LIB;;
esac
timeout 100 python bench.py --steps 2 --warmup 3 --T 40 --no-cpu-baseline --op-table gpurun_out/s12a_ops_lidc_$V.txt > gpurun_out/s12a_lidc_$V.json 2>&1
timeout 100 python bench.py --workload cityscapes --steps 2 --warmup 3 --T 20 --no-cpu-baseline --op-table gpurun_out/s12a_ops_cs_$V.txt > gpurun_out/s12a_cs_$V.json 2>&1
echo $V; grep -h "attention" gpurun_out/s12a_ops_lidc_$V.txt gpurun_out/s12a_ops_cs_$V.txt | cut -c1-60; tail -1 gpurun_out/s12a_ops_lidc_$V.txt; tail -1 gpurun_out/s12a_ops_cs_$V.txt
done
