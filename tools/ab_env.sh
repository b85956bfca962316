#!/bin/bash
# A/B of environment switches on the resident step: tools/ab_env.sh <rounds> "<ENV=.. ENV=..>" "<ENV=..>" ...
# Runs bench.py (no extras) once per variant and round, alternating the variants so box / clock drift hits all equally.
# Prints: variant | value | e2e | ms_per_step | conv roofline frac | SM MHz
rounds=$1; shift
for r in $(seq 1 "$rounds"); do
  for v in "$@"; do
    env $v python bench.py --no-extras --no-cpu-baseline --steps 40 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v |', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
  done
done
