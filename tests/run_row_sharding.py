"""torchrun worker: row-sharded rollout on WORLD_SIZE GPUs vs (a) the fp64 ORACLE evaluated on rank 0's GPU and (b) the
unsharded engine.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_row_sharding.py [H W B precision C graph|eager peer|nccl]
Prints one JSON line from rank 0 (max relative errors of the gathered latents, timing).  tests/test_gpu_row_sharding.py spawns
it and asserts the oracle error against the north-star tolerance."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sf_oracle as so  # noqa: E402
from streamingflow_b200.config import ode_cfg  # noqa: E402
from streamingflow_b200.layers.temporal_ode_bayes import NNFOwithBayesianJumps  # noqa: E402
from streamingflow_b200.row_sharding import RowShardedOde  # noqa: E402

H, W, B = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (96, 80, 2)))
precision = sys.argv[4] if len(sys.argv) > 4 else "bf16x3"
C = int(sys.argv[5]) if len(sys.argv) > 5 else 64
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
seed = 13
m = NNFOwithBayesianJumps(C, C, ode_cfg(C)).eval()
m.load_state_dict(so.recipe_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed, 1.0), strict=True)
m = m.to(dev)
m.precision = precision
times = [sorted([-1.0, -0.5, 0.0, -0.8, -0.6, -0.4, -0.2, 0.0])] * B
targets = [[-1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0]] * B
counts = [8] * B
g = torch.Generator().manual_seed(seed)
hx = torch.tanh(torch.randn(sum(counts), C, H, W, generator=g)).to(dev)
tape = torch.randn(18 * B, C, H, W, generator=g).to(dev)
use_graphs = not (len(sys.argv) > 6 and sys.argv[6] == "eager")
transport = sys.argv[7] if len(sys.argv) > 7 else None          # "peer" (NVLink peer-memory kernels, the default) | "nccl"
sharded = RowShardedOde(m, H, W, B, use_graphs=use_graphs, transport=transport)
with torch.no_grad():
    band, ro = sharded.integrate(hx, counts, times, targets, 0.05, noise=tape)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    band, ro = sharded.integrate(hx, counts, times, targets, 0.05, noise=tape)
    torch.cuda.synchronize()
    dist.barrier()
    dt_sharded = time.perf_counter() - t0
    full = sharded.gather_rows(band)
    if rank == 0:
        m._draw_noise = lambda n, h, w, device: tape[:max(n, 1)].contiguous()
        _, ref = m.integrate_latents(hx, counts, times, targets, 0.05)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, ref = m.integrate_latents(hx, counts, times, targets, 0.05)
        torch.cuda.synchronize()
        dt_single = time.perf_counter() - t0
        err = ((full.double() - ref.double()).abs().max() / ref.double().abs().max()).item()
        # the oracle (fp64, same ATen calls as the reference) on the same observations and noise tape, sample by sample
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        sd64 = {"g." + k: (v.double() if v.is_floating_point() else v) for k, v in m.state_dict().items()}
        it = iter(tape.double()[:, None])
        want = []
        for b in range(B):
            sch = so.build_schedule(times[b], targets[b], 0.05, True)
            _, path = so.integrate_latent(sd64, "g", hx[8 * b:8 * b + 8].double(), sch, it)
            want.append(torch.stack([path[i][0] for i in sch.select]))
        want = torch.stack(want)
        err_oracle = ((full.double() - want).abs().max() / want.abs().max()).item()
        err_single_oracle = ((ref.double() - want).abs().max() / want.abs().max()).item()
        print(json.dumps(dict(test="row_sharding", world=world, H=H, W=W, B=B, C=C, precision=precision, max_rel_err=err,
                              max_rel_err_vs_oracle=err_oracle, single_gpu_max_rel_err_vs_oracle=err_single_oracle,
                              ms_sharded=1e3 * dt_sharded, ms_single_gpu=1e3 * dt_single, events=len(ro.events),
                              halo_rows=12, band_rows=sharded.own_hi - sharded.own_lo,
                              launch=sharded.graph_mode, transport=sharded.transport, peer_error=sharded.peer_error, whole_graph_error=sharded.__dict__.get("_whole_graph_error"))), flush=True)
sharded.release_graphs()          # graphs with NCCL kernels must die before their process group; raises if a peer exchange timed out
dist.destroy_process_group()
