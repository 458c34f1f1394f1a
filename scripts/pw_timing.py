"""ConvNeXt block pointwise pair at B = 8, T = 7, 200x200x64: ONE back-to-back GEMM stage vs three launches (SF_PW_B2B=0)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from streamingflow_b200.refine_engine import RefineEngine
dev = torch.device("cuda", 0)
m = bench.make_model(dev)
B, T, C, H = 8, 7, 64, 200
sd = {k: v for k, v in m.state_dict().items() if k.startswith(("spatial_grus", "res_blocks"))}
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
x32 = torch.randn(B * T, H, H, C, device=dev)
pl = (x32.to(torch.bfloat16), None)
for fuse in ("0", "1"):
    os.environ["SF_PW_B2B"] = fuse
    eng = RefineEngine(sd, H, H, B, T, "bf16", dev)
    eng.run(pl, x32)
    t_pw = timed(lambda: eng.plan.run(eng.slots["block"], B * T))
    t_all = timed(lambda: eng.run_core(pl, x32), 5)
    print(f"SF_PW_B2B={fuse}: pointwise pair {t_pw:.3f} ms ({len(eng.slots['block'])} launches), whole refinement {t_all:.3f} ms", flush=True)
    del eng
