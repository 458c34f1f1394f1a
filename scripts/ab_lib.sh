#!/bin/bash
# A/B of several builds of libsf_b200.so on the same box: scripts/ab_lib.sh "build/lib_a.so build/lib_b.so ..." [reps]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
LIBS=$1; REPS=${2:-3}
cp streamingflow_b200/libsf_b200.so /tmp/lib_keep.so
for i in $(seq $REPS); do
  for L in $LIBS; do
    cp $L streamingflow_b200/libsf_b200.so
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-module-level --no-stage-timing 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$L', round(d['value'],1), 'ms', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
cp /tmp/lib_keep.so streamingflow_b200/libsf_b200.so
