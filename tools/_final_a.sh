set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r02f_gputests.log 2>&1; tail -3 gpurun_out/r02f_gputests.log
cp gpurun_out/parity_report.json gpurun_out/r02f_parity_report.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; tail -3 gpurun_out/r02f_smoke.log
python tools/chain_agreement.py lidc > gpurun_out/r02f_chain_agreement_lidc.json 2> gpurun_out/r02f_chain_lidc.err
python tools/chain_agreement.py cityscapes > gpurun_out/r02f_chain_agreement_cityscapes.json 2> gpurun_out/r02f_chain_cs.err
python bench.py --op-table gpurun_out/r02f_op_table_lidc_exact.txt > gpurun_out/r02f_bench_default.json 2> gpurun_out/r02f_bench_default.err
python bench.py --workload cityscapes --headline-only --no-cpu-baseline --op-table gpurun_out/r02f_op_table_cs_exact.txt > gpurun_out/r02f_bench_cs_exact.json 2>/dev/null
python bench.py --workload lidc --precision bf16 --headline-only --no-cpu-baseline --op-table gpurun_out/r02f_op_table_lidc_bf16.txt > gpurun_out/r02f_bench_lidc_bf16.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02f_bench_reference_arm.json 2>/dev/null
head -c 700 gpurun_out/r02f_bench_default.json
