"""Profiling driver for ncu: warm-up rollout, then ONE event (all stages) of the bench workload between
cudaProfilerStart/Stop.  Usage (on the GPU box):
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python scripts/profile_rollout.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_stage -o gpurun_out/prof \
      python scripts/profile_rollout.py
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", default="cell")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--events", type=int, default=1)
a = ap.parse_args()
hw = 200 if a.grid == "cell" else 50
dev = torch.device("cuda", 0)
model = bench.make_model(dev)
ode = model.gru_ode
ode.precision = a.precision
times = sorted(bench.CAM_T + bench.LIDAR_T)
hx = torch.tanh(torch.randn(a.batch * len(times), 64, hw, hw, device=dev))
with torch.no_grad():
    ode.integrate_latents(hx, [len(times)] * a.batch, [times] * a.batch, [bench.TARGETS] * a.batch, 0.05)
torch.cuda.synchronize()
eng = ode._engines[next(iter(ode._engines))]["engine"]
from streamingflow_b200 import engine as en  # noqa: E402

n = a.batch
ev = dict(kind=0, samples=list(range(n)), x_img=list(range(n)), rec=[-1] * n, eps=list(range(n)), dt=[0.1] * n, x_buf=en.BUF_X,
          s_in=0, s_base=0, s_out=0, run_cell=1, run_prior=1, want_f32=0)
table, evs = eng.build_table([ev] * a.events)
tdev = eng.upload_table(table)
eng.run_events(evs, tdev)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.run_events(evs, tdev)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", a.events, "event(s);", eng.lib.sf_plan_last_launches(eng.plan), "launches")
