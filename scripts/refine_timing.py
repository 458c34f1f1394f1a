import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0)
m = bench.make_model(dev)
B, T, C, H = 8, 7, 64, 200
x = torch.randn(B, T, C, H, H, device=dev)
def timed(fn, n=3):
    fn(); torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
with torch.no_grad():
    for i in range(2):
        print(f"spatial_gru[{i}] {timed(lambda: m.spatial_grus[i](x, x[:, 0])):.2f} ms")
        f = x.view(B * T, C, H, H)
        print(f"res_block[{i}] ({type(m.res_blocks[i]).__name__}) {timed(lambda: m.res_blocks[i](f)):.2f} ms")
    torch.backends.cudnn.benchmark = True
    for i in range(2):
        print(f"[cudnn.benchmark] spatial_gru[{i}] {timed(lambda: m.spatial_grus[i](x, x[:, 0])):.2f} ms  res_block {timed(lambda: m.res_blocks[i](x.view(B*T,C,H,H))):.2f} ms")
