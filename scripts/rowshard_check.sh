#!/bin/bash
# 2-GPU (or N-GPU) check of the row-sharded path: oracle parity in both graph modes + eager, then config-5 timing; short timeouts
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${N:-2}
run() { # mode args...
  MODE=$1; shift
  SF_ROWSHARD_GRAPH=$MODE timeout -k 5 ${T:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 295$((RANDOM % 90 + 10)) tests/run_row_sharding.py "$@" > gpurun_out/rowshard_chk.log 2>&1
  echo "exit $? ($MODE $*)"; grep "^{" gpurun_out/rowshard_chk.log | tee -a gpurun_out/rowshard_check.jsonl | cut -c1-330; grep -E "Error|error" gpurun_out/rowshard_chk.log | head -3
}
run segments 96 80 2 bf16x3 64 graph
run whole 96 80 2 bf16x3 64 graph
run segments 96 80 2 bf16 64 eager
run segments 64 48 1 bf16x3 128 graph
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29597 scripts/rowshard_timing.py whole segments 2>&1 | grep "^{" | tee gpurun_out/rowshard_timing_n$N.jsonl
