#!/bin/bash
# A/B inside ONE gpurun call (same box): usage tools/ab.sh "ENV_A" "ENV_B" [bench args]; alternates A B A B
A="$1"; B="$2"; shift 2
for i in 1 2; do
  for cfg in "$A" "$B"; do
    env $cfg python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'])"
  done
done
