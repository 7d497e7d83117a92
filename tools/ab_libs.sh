#!/bin/bash
# A/B of library builds on ONE box, exact mode only: tools/ab_libs.sh <lib1.so> <lib2.so> ...  (paths relative to ccdm_b200/; T=50 chains)
cd "$(dirname "$0")/.."
PKG=ccdm-stochastic-segmentation_b200/ccdm_b200
for rep in 1 2; do
for lib in "$@"; do
  for wl in lidc cityscapes; do
    CCDM_B200_LIB=$PWD/$PKG/$lib timeout 120 python bench.py --workload $wl --precision exact --headline-only --steps 2 --warmup 2 --T 50 --no-cpu-baseline --no-op-profile 2>/dev/null \
      | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$wl', 'ms/reverse-step %.4f' % (d['ms_per_step']/50))"
  done
done
done
