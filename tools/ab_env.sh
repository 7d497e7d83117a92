#!/bin/bash
# A/B of environment knobs on ONE box: tools/ab_env.sh "<workload> <precision>" "VAR=val ..." "VAR=val ..." ...   (short T=50 chains)
cd "$(dirname "$0")/.."
CFG=$1; shift
set -- "" "$@"
for rep in 1 2; do
  for envs in "$@"; do
    env $envs python bench.py --workload ${CFG% *} --precision ${CFG#* } --headline-only --steps 2 --warmup 2 --T 50 --no-cpu-baseline --no-op-profile 2>/dev/null \
      | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$CFG [$envs]', 'ms/reverse-step %.4f' % (d['ms_per_step']/50))"
  done
done
