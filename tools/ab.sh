#!/bin/bash
# A/B of library builds on ONE box: tools/ab.sh <lib1.so> <lib2.so> ...  (paths relative to ccdm_b200/); short chains (T=50), so
# the numbers are for comparison only.  Prints ms per reverse step per (lib, workload, precision).
cd "$(dirname "$0")/.."
PKG=ccdm-stochastic-segmentation_b200/ccdm_b200
LIBS="$@"
for rep in 1 2; do
for lib in $LIBS; do
  for cfg in "lidc exact" "lidc bf16" "cityscapes exact" "cityscapes bf16"; do
    set -- $cfg
    CCDM_B200_LIB=$PWD/$PKG/$lib python bench.py --workload $1 --precision $2 --headline-only --steps 2 --warmup 2 --T 50 --no-cpu-baseline --no-op-profile 2>/dev/null \
      | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$1', '$2', 'ms/reverse-step %.4f' % (d['ms_per_step']/50))"
  done
done
done
