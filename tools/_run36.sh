python -m pytest tests/test_gpu_variants.py -q 2>&1 | tail -6 > gpurun_out/r36_variants.log
for rep in 1 2 3; do
for lib in lib_base.so libccdm_b200.so; do
  for cfg in "lidc exact" "cityscapes exact"; do
    set -- $cfg
    CCDM_B200_LIB=$PWD/ccdm-stochastic-segmentation_b200/ccdm_b200/$lib python bench.py --workload $1 --precision $2 --headline-only --steps 2 --warmup 2 --T 50 --no-cpu-baseline --no-op-profile 2>/dev/null \
      | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$1', '$2', 'ms/reverse-step %.4f' % (d['ms_per_step']/50))"
  done
done
done > gpurun_out/r36_ab.log
cat gpurun_out/r36_variants.log gpurun_out/r36_ab.log
