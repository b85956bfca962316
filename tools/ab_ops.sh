#!/bin/bash
# A/B of environment switches on single launches of the B = 64 plan: tools/ab_ops.sh <rounds> <grep pattern> "<ENV=..>" "<ENV=..>" ...
# Runs tools/time_plan_ops.py once per variant and round (alternating) and prints the matching lines; launches timed alone see
# boost or capped clocks depending on the moment, so compare several rounds.
rounds=$1; pat=$2; shift 2
for r in $(seq 1 "$rounds"); do
  for v in "$@"; do
    echo "== $v (round $r)"
    env $v python tools/time_plan_ops.py 30 2>/dev/null | grep -E "$pat" | cut -c1-110
  done
done
