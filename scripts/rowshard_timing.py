"""torchrun worker: config 5 (400x400x128, B = 1) row-sharded over WORLD_SIZE GPUs; times whole-graph vs segment replay.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/rowshard_timing.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
c = bench.Ctx()
c.world, c.rank, c.local, c.dev = world, rank, local, dev


def barrier():
    dist.barrier()
    torch.cuda.synchronize()


c.barrier = barrier
for mode in (sys.argv[1:] or ["peer", "whole", "segments"]):          # peer = NVLink peer-memory kernels; whole / segments = NCCL graph modes
    os.environ["SF_ROWSHARD_TRANSPORT"] = "peer" if mode == "peer" else "nccl"
    os.environ["SF_ROWSHARD_GRAPH"] = mode
    out = bench.config5_row_sharded(c, steps=5)
    if rank == 0:
        print(json.dumps(dict(mode=mode, **{k: out.get(k) for k in ("value", "ms_per_rollout", "n_gpus", "launch", "transport", "peer_error",
                                                                    "whole_graph_error", "band_rows", "peer_trace", "phases_ms")})), flush=True)
dist.destroy_process_group()
