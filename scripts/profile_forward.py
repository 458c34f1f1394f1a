"""Profiling driver for ncu: ONE eager FuturePredictionODE.forward (B = 8, BEV 200x200x64, 8 observations, 7 targets: encoder,
step loop, decoder, refinement) between cudaProfilerStart/Stop -- the per-launch list of the e2e path.
  SF_B200_FORWARD_GRAPH=0 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/forward_launches.csv python scripts/profile_forward.py"""
import os
import sys

import torch

os.environ["SF_B200_FORWARD_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda", 0)
m = bench.make_model(dev)
B, H = int(os.environ.get("B", 8)), 200
with torch.no_grad():
    frames = torch.randn(B * 8, 64, H, H, device=dev)
    cam, lid = frames.view(B, 8, 64, H, H)[:, :3].contiguous(), frames.view(B, 8, 64, H, H)[:, 3:].contiguous()
    ct = torch.tensor([bench.CAM_T] * B, dtype=torch.float64)
    lt = torch.tensor([bench.LIDAR_T] * B, dtype=torch.float64)
    tt = torch.tensor([bench.TARGETS] * B, dtype=torch.float64)
    fpi = torch.zeros(B, 1, 64, H, H, device=dev)
    for _ in range(2):
        m(fpi, cam, lid, ct, lt, tt)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(fpi, cam, lid, ct, lt, tt)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled one forward")
