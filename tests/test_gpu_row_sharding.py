"""Row sharding over NCCL (BASELINE config 5) against the ORACLE: spawns a 2-rank torchrun of
tests/run_row_sharding.py (one process per GPU) and checks the gathered latents against the fp64 oracle rollout.
Needs two GPUs: the driver's single-GPU test run skips it; ``gpurun --gpus 2 -- python -m pytest tests/test_gpu_row_sharding.py``
runs it (log committed under profiles/)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, *argv):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "run_row_sharding.py")] + [str(a) for a in argv]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:]
    return json.loads(lines[-1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="row sharding needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("H,W,B,precision,C,launch", [(96, 80, 2, "bf16x3", 64, "graph"), (96, 80, 2, "bf16", 64, "graph"),
                                                      (64, 48, 1, "bf16x3", 128, "graph"), (96, 80, 2, "bf16x3", 64, "eager")])
def test_row_sharded_rollout_matches_oracle(H, W, B, precision, C, launch, transport):
    """Two ranks, each a band of rows + 12-row halos, halo exchange + SE all-reduce per event -- as kernels over NVLink peer
    memory ("peer": sf_halo_push / sf_halo_pull / sf_peer_allreduce_f32) or as NCCL calls -- full config-2 schedule (8 jumps + 10
    steps): the gathered selected latents vs the fp64 oracle, and vs the unsharded engine."""
    d = _run(2, H, W, B, precision, C, launch, transport)
    tol = 1e-2 if precision == "bf16" else 1e-4
    assert d["world"] == 2 and d["events"] == 18
    assert d["transport"] == transport, d          # no silent fall-back to NCCL when the peer arenas cannot be mapped
    assert d["max_rel_err_vs_oracle"] < tol, d
    assert d["single_gpu_max_rel_err_vs_oracle"] < tol, d
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "row_sharding_vs_oracle.jsonl"), "a") as f:
            f.write(json.dumps(d) + "\n")
