#!/bin/bash
# First-contact GPU run: bring-up tests in separate processes (a trap poisons the CUDA context), then parity, smoke, bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpu.txt
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -${TAILN:-25} gpurun_out/$name.log; }
run t1_bringup python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "device or tma or umma or layout" -x
run t2_event python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "one_event or batch_comp or zero_dt"
run t3_rollout python -m pytest tests/test_gpu_rollout.py -q -m gpu
run t4_smoke python -c "import __graft_entry__ as g; g.smoke()"
run t5_bench python bench.py --steps 3 --warmup 3
