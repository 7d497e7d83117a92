mkdir -p gpurun_out
timeout 60 python tools/tiny_chain.py bf16 128 8 > gpurun_out/s5a_tiny.log 2>&1; rc=$?; echo "tiny rc=$rc"; tail -1 gpurun_out/s5a_tiny.log
if [ $rc -eq 0 ]; then
timeout 240 python -m pytest tests -m gpu -q > gpurun_out/s5a_pytest.log 2>&1; tail -4 gpurun_out/s5a_pytest.log | cut -c1-250
for LN in 1 2 3 4 8; do
timeout 200 python bench.py --lanes $LN --steps 2 --warmup 3 --T 40 --no-cpu-baseline --no-op-profile > gpurun_out/s5a_lidc_$LN.json 2>&1; python -c "
import json
l=[x for x in open('gpurun_out/s5a_lidc_$LN.json') if x.startswith('{')][-1]; d=json.loads(l); print('lanes=$LN lidc ms/step', round(d['ms_per_step']/40,3), 'e2e', round(d['e2e']['ms_per_step']/40,3))"
done
for LN in 1 2 4 8; do
timeout 200 python bench.py --workload cityscapes --lanes $LN --steps 2 --warmup 3 --T 30 --no-cpu-baseline --no-op-profile > gpurun_out/s5a_cs_$LN.json 2>&1; python -c "
import json
l=[x for x in open('gpurun_out/s5a_cs_$LN.json') if x.startswith('{')][-1]; d=json.loads(l); print('lanes=$LN cs ms/step', round(d['ms_per_step']/30,3))"
done
fi
