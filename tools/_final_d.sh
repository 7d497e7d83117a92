python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 > gpurun_out/r02f_bench_8gpu.json 2> gpurun_out/r02f_bench_8gpu.err
grep '^{' gpurun_out/r02f_bench_8gpu.json | head -c 300
