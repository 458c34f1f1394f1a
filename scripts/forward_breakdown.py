"""Where FuturePredictionODE.forward (B = 8, BEV 200x200x64, 8 observations, 7 targets) spends its time on the engine:
encoder / ODE loop / decoder / refinement (per part), plus the BEV Decoder head.  CUDA events, device-resident inputs."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from streamingflow_b200 import _lib as L  # noqa: E402
from streamingflow_b200 import refine_engine as rf  # noqa: E402

dev = torch.device("cuda", 0)
m = bench.make_model(dev)
ode = m.gru_ode
B, H = int(os.environ.get("B", 8)), 200
times = sorted(bench.CAM_T + bench.LIDAR_T)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


with torch.no_grad():
    frames = torch.randn(B * 8, 64, H, H, device=dev)
    cam, lid = frames.view(B, 8, 64, H, H)[:, :3].contiguous(), frames.view(B, 8, 64, H, H)[:, 3:].contiguous()
    ct = torch.tensor([bench.CAM_T] * B, dtype=torch.float64)
    lt = torch.tensor([bench.LIDAR_T] * B, dtype=torch.float64)
    tt = torch.tensor([bench.TARGETS] * B, dtype=torch.float64)
    fpi = torch.zeros(B, 1, 64, H, H, device=dev)
    t_full = timed(lambda: m(fpi, cam, lid, ct, lt, tt))
    codec = ode._codec_for(H, H, B * 8, B * 7, dev)
    t_enc = timed(lambda: codec.encode(frames))
    planes = codec.encode(frames)
    run_ode = lambda: ode._integrate_impl(None, [8] * B, [times] * B, [bench.TARGETS] * B, 0.05, obs_planes=planes, return_slots=True)
    t_ode = timed(run_ode)
    _, (eng, flat) = run_ode()
    slots = torch.tensor(flat, dtype=torch.int32).to(dev)
    t_dec = timed(lambda: codec.decode(eng.path, slots, unpack=False))
    pl, x32 = codec.decode(eng.path, slots, unpack=False)
    refine = m._refine_for(H, H, B, 7, dev)
    t_ref = timed(lambda: refine.run(pl, x32))
    # inside the refinement
    P, lib, n = refine.plan, refine.lib, B * 7
    stream = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    parts = {}
    P.buf(rf.R_X, 64, planes=(pl[0][:n], pl[1][:n] if pl[1] is not None else None))
    parts["init_state"] = timed(lambda: refine._init_state(pl, x32))
    parts["spatial_gru0 (7 x gates+propose+1x1)"] = timed(lambda: refine._run_gru(0, pl))
    blk = refine.g["block"]
    o0, dw = P.bufs[rf.R_O0], P.bufs[rf.R_DW]
    ptr = lambda t: t.data_ptr() if t is not None else None
    parts["block.dwconv7+LN"] = timed(lambda: L.check(lib.sf_dwconv7_ln(ptr(o0[0]), ptr(o0[1]), ptr(dw[0]), ptr(dw[1]), blk["dw_w"].data_ptr(), blk["dw_b"].data_ptr(),
                                                                      blk["ln_w"].data_ptr(), blk["ln_b"].data_ptr(), n, 64, H, H, stream()), "dw"))
    parts["block.pw1+pw2"] = timed(lambda: P.run(refine.slots["block"], n))
    parts["spatial_gru1"] = timed(lambda: refine._run_gru(1, pl))
    names = [s.name for s in refine.g["deeplab"]]
    for slot, name in zip(refine.slots["deeplab"], names):
        parts["deeplab." + name] = timed(lambda s=slot: P.run([s], n))
    out = torch.empty((n, 64, H, H), dtype=torch.float32, device=dev)
    parts["unpack NHWC->NCHW fp32"] = timed(lambda: L.check(lib.sf_unpack_nhwc_f32(refine.out32.data_ptr(), out.data_ptr(), None, n, 64, H, H, stream()), "unpack"))
    dec = bench.make_decoder(dev)
    x = m(fpi, cam, lid, ct, lt, tt)[0]
    t_head = timed(lambda: dec(x, planes=m.last_output_planes))
    t_head_nchw = timed(lambda: dec(x))
ro = ode.last_rollout
print(f"B={B} forward {t_full:.2f} ms = {ro.n_state_steps / t_full * 1e3:.0f} state-steps/s | encoder {t_enc:.2f}  ode-loop {t_ode:.2f}  decoder {t_dec:.2f}  "
      f"refinement {t_ref:.2f} | Decoder head (planes) {t_head:.2f}  (NCHW fp32 in) {t_head_nchw:.2f} ms")
for k, v in parts.items():
    print(f"    refine.{k:42s} {v:7.3f} ms")
