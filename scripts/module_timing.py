"""Times FuturePredictionODE.forward (module level: BEV 200x200x64, B=8, 8 obs, 7 targets) and its parts on the GPU."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0)
m = bench.make_model(dev)
B, C, H = int(os.environ.get("B", 8)), 64, 200
cam = torch.randn(B, 3, C, H, H, device=dev); lid = torch.randn(B, 5, C, H, H, device=dev)
ct = torch.tensor([bench.CAM_T] * B, dtype=torch.float64); lt = torch.tensor([bench.LIDAR_T] * B, dtype=torch.float64)
tt = torch.tensor([bench.TARGETS] * B, dtype=torch.float64)
fpi = torch.zeros(B, 1, C, H, H, device=dev)
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) * 1e3 / n
with torch.no_grad():
    full = timed(lambda: m(fpi, cam, lid, ct, lt, tt))
    frames = torch.cat([cam, lid], 1).reshape(B * 8, C, H, H)
    enc = timed(lambda: m.gru_ode.srvp_encoder(frames))
    hx = m.gru_ode.srvp_encoder(frames)
    times = sorted(bench.CAM_T + bench.LIDAR_T)
    ode = timed(lambda: m.gru_ode.integrate_latents(hx, [8] * B, [times] * B, [bench.TARGETS] * B, 0.05))
    sel = m.gru_ode.integrate_latents(hx, [8] * B, [times] * B, [bench.TARGETS] * B, 0.05)[1]
    dec = timed(lambda: m.gru_ode.srvp_decode(sel))
    x = m.gru_ode.srvp_decode(sel)
    def refine():
        y = x; h0 = y[:, 0]
        for g, blk in zip(m.spatial_grus, m.res_blocks):
            y = g(y, h0); b, s, c, h, w = y.shape; y = blk(y.view(b * s, c, h, w)).view(b, s, c, h, w)
        return y
    ref = timed(refine)
ro = m.gru_ode.last_rollout
print(f"B={B} module forward: {full[0]:.2f} ms GPU ({full[1]:.2f} ms wall) -> {ro.n_state_steps/full[0]*1e3:.0f} state-steps/s")
print(f"  encoder {enc[0]:.2f}  ode-loop {ode[0]:.2f} (wall {ode[1]:.2f}; {ro.n_state_steps/ode[0]*1e3:.0f} state-steps/s; launches {ro.launches})  decoder {dec[0]:.2f}  refinement {ref[0]:.2f} ms")
