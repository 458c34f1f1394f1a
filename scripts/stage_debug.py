"""Experiment: time each conv stage alone with the epilogues (SF_DEBUG_STAGE=1) or the MMAs (=2) switched off."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0)
model = bench.make_model(dev)
ode = model.gru_ode
times = sorted(bench.CAM_T + bench.LIDAR_T)
B, hw = 8, 200
hx = torch.tanh(torch.randn(B * len(times), 64, hw, hw, device=dev))
with torch.no_grad():
    ode.integrate_latents(hx, [len(times)] * B, [times] * B, [bench.TARGETS] * B, 0.05)
eng = ode._engines[next(iter(ode._engines))]["engine"]
st = bench.time_stages(eng, B, hw, bench.load_peaks())
print(os.environ.get("SF_DEBUG_STAGE", "0"), " ".join(f"{k}:{v['ms_per_event']*1e3:.0f}" for k, v in st.items()))
