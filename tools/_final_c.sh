python -m pytest tests/test_gpu_sharded.py -q > gpurun_out/r02f_sharded_tests.log 2>&1; tail -2 gpurun_out/r02f_sharded_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02f_bench_2gpu.json 2> gpurun_out/r02f_bench_2gpu.err
grep '^{' gpurun_out/r02f_bench_2gpu.json | head -c 400
