mkdir -p gpurun_out
for M in 1 0 1 0; do
CCDM_IDENT_SKIP=$M timeout 60 python tools/tiny_chain.py bf16 128 8 > gpurun_out/s9c_tiny.log 2>&1 || { echo "tiny chain FAILED"; tail -5 gpurun_out/s9c_tiny.log; exit 1; }
CCDM_IDENT_SKIP=$M timeout 100 python bench.py --steps 2 --warmup 3 --T 40 --no-cpu-baseline --no-op-profile > gpurun_out/s9c_lidc_$M.json 2>&1
CCDM_IDENT_SKIP=$M timeout 100 python bench.py --workload cityscapes --steps 2 --warmup 3 --T 20 --no-cpu-baseline --no-op-profile > gpurun_out/s9c_cs_$M.json 2>&1
echo "IDENT_SKIP=$M"; python - <<PY
import json
for w in ("lidc","cs"):
    try:
        d=json.loads(open(f"gpurun_out/s9c_{w}_$M.json").read().strip().splitlines()[-1]); print(w, d["value"], d["ms_per_step"])
    except Exception as e: print(w, "ERR", e)
PY
done
