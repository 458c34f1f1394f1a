#!/bin/bash
# A/B of an environment toggle on the same box: alternates the two settings, device-resident rollout only.
# usage: scripts/ab_bench.sh VAR [reps]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
VAR=$1; REPS=${2:-3}
for i in $(seq $REPS); do
  for v in 1 0; do
    env $VAR=$v python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-module-level --no-stage-timing 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
