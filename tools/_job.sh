mkdir -p gpurun_out
PREV=$PWD/ccdm-stochastic-segmentation_b200/ccdm_b200/libccdm_b200_prev.so
timeout 60 python tools/tiny_chain.py bf16 128 8 > gpurun_out/s10c_tiny.log 2>&1 || { echo "tiny chain FAILED"; tail -5 gpurun_out/s10c_tiny.log; exit 1; }
for V in prev new0 new1 prev new0 new1; do
case $V in
prev) export CCDM_B200_LIB=$PREV; export CCDM_STREAM=0;;
new0) unset CCDM_B200_LIB; export CCDM_STREAM=0;;
new1) unset CCDM_B200_LIB; export CCDM_STREAM=1;;
esac
timeout 100 python bench.py --steps 2 --warmup 3 --T 40 --no-cpu-baseline --no-op-profile > gpurun_out/s10c_lidc_$V.json 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/s10c_lidc_$V.json").read().strip().splitlines()[-1]); print("$V", d["ms_per_step"])
PY
done
