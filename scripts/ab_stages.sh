#!/bin/bash
# A/B with the stage table: scripts/ab_stages.sh "VAR=a,VAR2=b VAR=c,VAR2=d"  (one bench --quick run per setting, stage timing on)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for S in $1; do
  env ${S//,/ } python bench.py --steps 8 --warmup 3 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$S', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'fwd', round(d['e2e'].get('forward_ms_in_pipeline',0),2), d['clocks']['sm_mhz'], 'parity', d['parity'].get('max_rel_err'))
print('   ', ' '.join(f\"{k}={v['ms_per_event']*1e3:.1f}\" for k,v in (d.get('stages') or {}).items()))"
done
