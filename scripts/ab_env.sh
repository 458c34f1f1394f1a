#!/bin/bash
# A/B of environment switches on the same box: scripts/ab_env.sh "SF_B2B=0 SF_B2B=1" [reps]   (each token = one setting; VAR=a,VAR2=b sets two)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
SETTINGS=$1; REPS=${2:-3}
for i in $(seq $REPS); do
  for S in $SETTINGS; do
    env ${S//,/ } python bench.py --steps 8 --warmup 3 --quick --no-stage-timing 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$S', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'], 'parity', d['parity'].get('max_rel_err'))"
  done
done
