mkdir -p gpurun_out
PREV=$PWD/ccdm-stochastic-segmentation_b200/ccdm_b200/libccdm_b200_prev.so
timeout 60 python tools/tiny_chain.py bf16 128 8 > gpurun_out/s11d_tiny.log 2>&1 || { echo "tiny chain FAILED"; tail -5 gpurun_out/s11d_tiny.log; exit 1; }
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/s11d_pytest.log 2>&1; tail -2 gpurun_out/s11d_pytest.log | cut -c1-200
for V in prev new prev new; do
case $V in
prev) export CCDM_B200_LIB=$PREV;;
new) unset CCDM_B200_LIB;;
esac
timeout 100 python bench.py --steps 2 --warmup 3 --T 40 --no-cpu-baseline --no-op-profile > gpurun_out/s11d_lidc_$V.json 2>&1
timeout 100 python bench.py --workload cityscapes --steps 2 --warmup 3 --T 20 --no-cpu-baseline --no-op-profile > gpurun_out/s11d_cs_$V.json 2>&1
python - <<PY
import json
for w in ("lidc","cs"):
    d=json.loads(open("gpurun_out/s11d_%s_$V.json" % w).read().strip().splitlines()[-1]); print("$V", w, d["ms_per_step"])
PY
done
