"""ONE GPU: what a rank of the row-sharded config-5 rollout (400 columns x 128 channels, B = 1) spends in its own kernels, without
any collective -- the local image of `rows` rows integrated by RowShardedOde(world = 1).  Separates the compute of band + halos
(tile / wave quantisation included) from the exchange chain when read next to scripts/rowshard_timing.py.
    python scripts/rowshard_local_cost.py 400 212 124 74 64"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from streamingflow_b200.config import ode_cfg  # noqa: E402
from streamingflow_b200.layers.temporal_ode_bayes import NNFOwithBayesianJumps  # noqa: E402
from streamingflow_b200.row_sharding import RowShardedOde  # noqa: E402

C, W = 128, 400
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = NNFOwithBayesianJumps(C, C, ode_cfg(C)).eval().to(dev)
m.precision = "bf16"
times = sorted(bench.CAM_T + bench.LIDAR_T)
for rows in [int(a) for a in (sys.argv[1:] or ["400", "212", "124", "74", "64"])]:
    g = torch.Generator(device=dev).manual_seed(5)
    hx = torch.tanh(torch.randn(len(times), C, rows, W, device=dev, generator=g))
    tape = torch.randn(18, C, rows, W, device=dev, generator=g)
    sh = RowShardedOde(m, rows, W, 1)

    def rollout():
        with torch.no_grad():
            return sh.integrate(hx, [len(times)], [times], [bench.TARGETS], 0.05, noise=tape)

    rollout(); rollout()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        _, ro = rollout()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    n_ev = ro.n_state_steps + ro.n_jumps
    # per-stage time of one event's launches, eager, each launch timed alone (10 repetitions)
    eng = sh.eng
    evs = sh._tables[next(iter(sh._tables))][0]
    tdev = sh._tables[next(iter(sh._tables))][1]
    ev = evs[1]
    per = {}
    for op in sh._event_ops(ev):
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            sh._run_op(op, ev, tdev)
        b.record()
        torch.cuda.synchronize()
        per[f"{op[0]}:{op[1]}"] = round(a.elapsed_time(b) / 10 * 1e3, 1)
    print(json.dumps(dict(rows=rows, ms_per_rollout=round(ms, 3), events=n_ev, us_per_event=round(ms / n_ev * 1e3, 1),
                          launches_per_rollout=sh.launches // 7, stage_us=per, stage_sum_us=round(sum(per.values()), 1))), flush=True)
    sh.release_graphs()
    del sh
