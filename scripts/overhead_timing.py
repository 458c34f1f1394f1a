"""Times the non-stage parts of one bench rollout (B = 8, 200x200x64): noise draws, layout pack, gathers, graph replay."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

dev = torch.device("cuda", 0)
model = bench.make_model(dev)
ode = model.gru_ode
ode.cuda_graph = True
times = sorted(bench.CAM_T + bench.LIDAR_T)
B, hw = 8, 200
hx = torch.tanh(torch.randn(B * len(times), 64, hw, hw, device=dev))

def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def rollout():
    with torch.no_grad():
        return ode.integrate_latents(hx, [len(times)] * B, [times] * B, [bench.TARGETS] * B, 0.05)

print("rollout ms", timed(rollout))
eng = ode._engines[next(iter(ode._engines))]["engine"]
ent = next(iter(ode._graphs.values()))
print("graph replay ms", timed(lambda: ent["graph"].replay()))
print("pack ms", timed(lambda: eng.pack_into(3, hx)))
print("gather path ms", timed(lambda: eng.unpack_path(ent["slots"])))
print("gather final ms", timed(lambda: eng.unpack_f32(eng.state32[0], B)))
eps = ent["eps"]
n = eps.shape[0]
print("noise slots", n)
def per_slot():
    for i in range(n):
        eps[i].normal_()
print("noise per-slot ms", timed(per_slot))
print("noise bulk ms", timed(lambda: eps.normal_()))
print("zero state ms", timed(lambda: eng.zero_state(0)))
print("noise fused (one launch, reference stream) ms", timed(lambda: ode._draw_noise(n, hw, hw, dev, out=eps)))
