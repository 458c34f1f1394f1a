#!/bin/bash
# A/B of an environment toggle on the same box: alternates the settings, device-resident rollout only.
# usage: scripts/ab_bench.sh VAR "v1 v2 ..." [reps]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
VAR=$1; VALS=${2:-"1 0"}; REPS=${3:-3}
for i in $(seq $REPS); do
  for v in $VALS; do
    env $VAR=$v python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-module-level --no-stage-timing 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['gpu_launches'])"
  done
done
