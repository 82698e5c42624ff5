#!/bin/bash
# A/B of environment-selected kernel variants inside ONE gpurun call (boxes differ by a few percent, so only
# numbers from the same call are comparable).  Usage: tools/ab.sh "VAR=a" "VAR=b" ...   ("-" = no override)
for cfg in "$@"; do
  if [ "$cfg" = "-" ]; then envs=""; else envs="$cfg"; fi
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err || tail -3 gpurun_out/ab_tmp.err
  python - "$cfg" <<'PY'
import json, sys
d = json.load(open('gpurun_out/ab_tmp.json'))
b = d['breakdown']
print(f"{sys.argv[1]:32s} step {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_step']:.3f}  enc {b['encoders_ms']:.3f}  dec {b['decoder_ms']:.3f}  iter {b['per_iter_ms']:.4f}  zr {d['roofline']['ms_per_launch']*1e3:.1f} us  clk {d['clocks']['sm_mhz']}")
PY
done
