#!/bin/bash
# 2-GPU check of the row-sharded worker in both graph modes, each under its own short timeout (a hang must not eat the budget)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for MODE in segments whole; do
  echo "== $MODE"
  SF_ROWSHARD_GRAPH=$MODE timeout -k 5 ${T:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 2951$((RANDOM % 10)) tests/run_row_sharding.py ${ARGS:-96 80 2 bf16x3 64 graph} > gpurun_out/rowshard_dbg_$MODE.log 2>&1
  echo "exit $?"; grep "^{" gpurun_out/rowshard_dbg_$MODE.log | cut -c1-400; tail -3 gpurun_out/rowshard_dbg_$MODE.log | cut -c1-300
done
