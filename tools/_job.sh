TAG=r01g
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --op-table gpurun_out/${TAG}_ops_lidc.txt > gpurun_out/${TAG}_bench_lidc.json 2> gpurun_out/${TAG}_bench_lidc.err; tail -c 600 gpurun_out/${TAG}_bench_lidc.json; tail -3 gpurun_out/${TAG}_bench_lidc.err
timeout 900 python bench.py --workload cityscapes --steps 2 --warmup 3 --cpu-budget 10 --op-table gpurun_out/${TAG}_ops_cs.txt > gpurun_out/${TAG}_bench_cs.json 2> gpurun_out/${TAG}_bench_cs.err; tail -c 600 gpurun_out/${TAG}_bench_cs.json; tail -3 gpurun_out/${TAG}_bench_cs.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
