#!/bin/bash
rounds=$1; shift
for r in $(seq 1 "$rounds"); do
  for v in "$@"; do
    python bench.py --no-extras --no-cpu-baseline $v 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v |', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['sm_min_mhz'])"
  done
done
