#!/bin/bash
# A/B the bench between library dirs / env settings on one box.
#   tools/ab.sh "<label>|<env assignments>" ...      (batches from $BATCHES, default "8 1 32")
BATCHES=${BATCHES:-"8 1 32"}
STEPS=${STEPS:-200}
for spec in "$@"; do
  label=${spec%%|*}; envs=${spec#*|}
  for b in $BATCHES; do
    line=$(env $envs timeout 300 python bench.py --steps $STEPS --warmup 10 --batch $b --no-cpu-baseline ${BENCH_ARGS} 2>&1 | tail -1)
    echo "$line" | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    e = d['e2e']; ser = e.get('serial', e)
    print('%-28s b%-3d %8.0f img/s  %7.1f us | e2e %8.0f img/s %7.1f us (serial %7.1f us)' % ('$label', $b, d['value'], d['ms_per_step']*1e3, e['value'], e['ms_per_step']*1e3, ser['ms_per_step']*1e3))
except Exception as e:
    print('$label', $b, 'FAILED', e)
"
  done
done
