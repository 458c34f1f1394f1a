#!/bin/bash
# 8-GPU check of the row-sharded path: oracle parity (peer transport, 16-row bands) + config-5 timing, peer vs NCCL on the same box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${N:-8}
run() {
  timeout -k 5 ${T:-120} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 295$((RANDOM % 90 + 10)) tests/run_row_sharding.py "$@" > gpurun_out/rowshard_chk.log 2>&1
  echo "exit $? ($*)"; grep "^{" gpurun_out/rowshard_chk.log | tee -a gpurun_out/rowshard_check_n$N.jsonl | cut -c1-420; grep -E "Error|error" gpurun_out/rowshard_chk.log | grep -v '"' | head -3
}
run 128 80 2 bf16x3 64 graph peer
run 128 48 1 bf16 128 graph peer
timeout -k 5 ${TT:-200} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29597 scripts/rowshard_timing.py ${MODES:-peer whole} 2>&1 | grep "^{" | tee gpurun_out/rowshard_timing_n$N.jsonl
