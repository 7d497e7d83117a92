#!/bin/bash
# A/B of environment knobs at a given per-GPU batch: tools/ab_batch.sh "<workload> <precision> <batch>" "VAR=val ..." ...   (T=50 chains)
cd "$(dirname "$0")/.."
set -- $1 "${@:2}"
WL=$1; PREC=$2; B=$3; shift 3
for rep in 1 2; do
  for envs in "" "$@"; do
    env $envs python bench.py --workload $WL --precision $PREC --batch $B --headline-only --steps 2 --warmup 2 --T 50 --no-cpu-baseline --no-op-profile 2>/dev/null \
      | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$WL $PREC B=$B [$envs]', 'ms/reverse-step %.4f' % (d['ms_per_step']/50))"
  done
done
