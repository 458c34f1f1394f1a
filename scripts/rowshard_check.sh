#!/bin/bash
# 2-GPU (or N-GPU) check of the row-sharded path: oracle parity with the peer-memory transport (graph + eager, C = 64 / 128) and one
# NCCL case per graph mode, then config-5 timing of the three launch modes; every multi-rank command under its own short timeout
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${N:-2}
run() { # graph-mode args...
  MODE=$1; shift
  SF_ROWSHARD_GRAPH=$MODE timeout -k 5 ${T:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 295$((RANDOM % 90 + 10)) tests/run_row_sharding.py "$@" > gpurun_out/rowshard_chk.log 2>&1
  echo "exit $? ($MODE $*)"; grep "^{" gpurun_out/rowshard_chk.log | tee -a gpurun_out/rowshard_check.jsonl | cut -c1-420; grep -E "Error|error" gpurun_out/rowshard_chk.log | grep -v '"' | head -3
}
run segments 96 80 2 bf16x3 64 graph peer
run segments 96 80 2 bf16 64 eager peer
run segments 64 48 1 bf16x3 128 graph peer
if [ -z "$SKIP_NCCL" ]; then
run segments 96 80 2 bf16x3 64 graph nccl
run whole 96 80 2 bf16x3 64 graph nccl
fi
timeout -k 5 ${TT:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29597 scripts/rowshard_timing.py ${MODES:-peer whole segments} 2>&1 | grep "^{" | tee gpurun_out/rowshard_timing_n$N.jsonl
