#!/usr/bin/env python
"""bench.py -- ODE BEV state-steps/s of the GRU-ODE-Bayes integration path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid cell|module] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "config 2" of SURVEY.md 8d): B = 8 samples per GPU, camera 2 Hz (-1, -0.5, 0) +
LiDAR 5 Hz (-0.8 .. 0) observations, 7 targets (-1 .. 2 s), variable-step Euler, IMPUTE: 8 observation jumps + 10 ODE
state-steps per sample.  --grid cell (default) integrates the literal 200x200x64 state the metric names; --grid module
integrates the 50x50x64 latent the reference's module produces from a 200x200x64 BEV.
A "step" of the bench = one full rollout over the batch.  state-step := one sample x one ode_step call.

  value   device-resident: encoded observations already in HBM; timed: layout pack, noise draw, every stage kernel,
          path gather.  CUDA events, max over ranks.
  e2e     the same rollout through the public latent-level call with HOST (pinned) buffers: H2D of the observations
          and D2H of the selected states inside the timed region.
  roofline  per conv stage: algorithmic FLOPs (reference MACs x 2, no credit for zero padding) / CUDA-event time of that
          stage's launches on the launching stream, vs the measured bf16 GEMM peak (MEASURED_PEAKS.json).
  cpu_baseline / --impl reference: the oracle port of the reference algorithm (same torch ATen calls) on the host cores,
          bounded sample = 1 sample of the batch, full schedule.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CAM_T = [-1.0, -0.5, 0.0]
LIDAR_T = [-0.8, -0.6, -0.4, -0.2, 0.0]
TARGETS = [-1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0]
METRIC = "ode_bev_state_steps_per_sec"
UNIT = "state-steps/s"
# reference MACs per output pixel of each conv stage at C = 64 (cell 227 C^2 + 2C, p_model 137 C^2)
STAGE_MACS = {"gates": 4 * 9 * 128 * 64, "propose": 2 * 9 * 128 * 64, "decode": 9 * 64 * 64, "trunk7": 49 * 128 * 64,
              "trunk1": 64 * 64, "mix": 9 * 64 * 64 + 128 * 64 + 2 * 64, "q1": 9 * 64 * 64, "q2": 9 * 64 * 128 + 64 * 128,
              "q3": 9 * 128 * 128, "q4": 9 * 128 * 128, "q5": 9 * 128 * 128}
FLOPS_PER_STATE_STEP_PX = 2 * (364 * 64 * 64 + 2 * 64)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), hbm=d["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc, self.stop, self.thread = gpu_index, [], None, threading.Event(), None

    def __enter__(self):
        # NVML polled from a thread every 10 ms (the timed region is ~0.1 s; nvidia-smi -lms often delivers no sample in time)
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)

            def poll():
                while not self.stop.is_set():
                    try:
                        r = get_reasons(h)
                        self.rows.append([str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(mx), "0"] +
                                         ["Active" if r & m else "Not Active" for _, m in bits])
                    except Exception:
                        pass
                    self.stop.wait(0.01)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return self
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        elif self.thread is not None:
            self.thread.join(timeout=1)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def make_model(device):
    import torch
    from streamingflow_b200.config import ode_cfg
    from streamingflow_b200.models.future_prediction_ode import FuturePredictionODE

    torch.manual_seed(0)
    m = FuturePredictionODE(64, 64, 4, ode_cfg(64)).eval()     # random-init weights, torch default init (seed 0)
    return m.to(device)


def run_reference_arm(args):
    """The reference algorithm's CPU implementation (oracle port: the same torch ATen calls as the reference's nn.Modules),
    all host threads, on a bounded sample of the workload: ONE sample of the batch, full schedule, per step."""
    import torch
    from oracle import sf_oracle as so

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    hw = 200 if args.grid == "cell" else 50
    m = make_model("cpu")
    sd = {"g." + k: v for k, v in m.gru_ode.state_dict().items()}
    times = sorted(CAM_T + LIDAR_T)
    sch = so.build_schedule(times, TARGETS, 0.05, True)
    n_steps = sum(1 for e in sch.events if e.kind == "step")
    g = torch.Generator().manual_seed(1)
    hx = torch.tanh(torch.randn(len(times), 64, hw, hw, generator=g))

    def one():
        eps = (torch.randn(1, 64, hw, hw, generator=g) for _ in range(10 ** 6))
        with torch.no_grad():
            so.integrate_latent(sd, "g", hx, sch, eps)

    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    value = n_steps * args.steps / dt
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=workload_config(args, hw, 1),
                cpu_baseline=dict(value=value, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                                  sample=f"1 sample of the batch, full schedule ({len(sch.events)} events, {n_steps} state-steps) per step, fp32"),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def workload_config(args, hw, batch):
    return dict(workload=f"config2: Prediction_LC_ODE_Variable ODE head, batch {batch}/GPU, camera 2Hz + LiDAR 5Hz synthetic BEV "
                         f"observations (8 jumps + 10 variable Euler state-steps per sample), ODE grid {hw}x{hw}x64 "
                         f"({'cell-level: the literal 200x200x64 state' if args.grid == 'cell' else 'module-level latent of a 200x200x64 BEV'})",
                grid=args.grid, batch_per_gpu=batch, precision=args.precision, solver="euler", variable_step=True, impute=True,
                launch="eager" if getattr(args, "no_graph", True) else "one CUDA graph per rollout (all stage launches); layout pack, one-launch noise draw and gathers eager around it",
                l2="inputs + workspace (>1 GB at 200x200, B=8) exceed the 126 MB L2; no explicit flush" if hw >= 200 else
                   "working set fits L2 (module-level latent): L2 flushed by a 256 MB memset between steps",
                parallelism=f"batch-sharded x{args.gpus}, no collective in the data path")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", default="cell", choices=["cell", "module"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stage-timing", action="store_true")
    ap.add_argument("--no-module-level", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the stages eagerly instead of replaying the captured CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warm = max(args.warmup, 3)

    hw = 200 if args.grid == "cell" else 50
    B = args.batch
    model = make_model(dev)
    ode = model.gru_ode
    ode.precision = args.precision
    ode.cuda_graph = not args.no_graph       # the whole step loop is one captured CUDA graph (eager launches with --no-graph)
    times = sorted(CAM_T + LIDAR_T)
    n_obs = len(times)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    hx_dev = torch.tanh(torch.randn(B * n_obs, 64, hw, hw, device=dev, generator=g))      # encoded observations (tanh head range)
    hx_host = hx_dev.cpu().pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if hw < 200 else None

    def rollout(hx):
        with torch.no_grad():
            return ode.integrate_latents(hx, [n_obs] * B, [times] * B, [TARGETS] * B, 0.05)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing
    for _ in range(warm):
        rollout(hx_dev)
    eng = ode._engines[next(iter(ode._engines))]["engine"]
    ro = ode.last_rollout
    steps_per_rollout = ro.n_state_steps
    barrier()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        barrier()
        for a, b in ev:
            if flush is not None:
                flush.zero_()
            a.record()
            rollout(hx_dev)
            b.record()
        barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = ro.launches + 4               # stage kernels of one rollout + layout pack + noise fill + path gather + final-state gather
    eng.check_errflag()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    value = world * steps_per_rollout * args.steps / (ms_total * 1e-3)

    # ---------------- end to end with host buffers
    sel_host = torch.empty((len(TARGETS), B, 64, hw, hw), dtype=torch.float32).pin_memory()

    def e2e_once():
        with torch.no_grad():
            # join=False: the downloads of this step stay on their copy stream; the device-wide synchronize that closes the
            # timed region waits for them, and the next step's uploads / compute pipeline behind this one
            ode.integrate_latents_streamed(hx_host, [n_obs] * B, [times] * B, [TARGETS] * B, 0.05, out_host=sel_host, join=False)

    e2e_once()
    barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(2, min(args.steps, 5))
    ea.record()
    for _ in range(n_e2e):
        e2e_once()
    torch.cuda.current_stream().wait_event(ode.download_done)      # the closing event comes after the last step's last download
    eb.record()
    barrier()
    t = torch.tensor([ea.elapsed_time(eb)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * steps_per_rollout * n_e2e / (t.item() * 1e-3)
    e2e = dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=hx_host.numel() * 4, d2h_bytes_per_step=sel_host.numel() * 4,
               api="NNFOwithBayesianJumps.integrate_latents_streamed (the body of forward between srvp_encode and srvp_decode) on pinned "
                   "host buffers; uploads / downloads pipelined against the rollout on copy streams (double-buffered staging: consecutive steps "
                   "pipeline into each other), all copies of every timed step inside the timed region")

    # ---------------- per-stage roofline (rank 0)
    peaks = load_peaks()
    roof, stages = None, None
    if rank == 0 and not args.no_stage_timing:
        stages = time_stages(eng, B, hw, peaks)
        total = sum(s["ms_per_event"] for s in stages.values())
        dom = max((k for k in stages if "tflops" in stages[k]), key=lambda k: stages[k]["ms_per_event"])
        d = stages[dom]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(f"{dom}@{hw}x{B}")
        roof = dict(bound="tensor", kernel=f"conv_stage_kernel<{dom}>", achieved=d["tflops"], peak=peaks["bf16"], unit="TFLOP/s",
                    frac=d["tflops"] / peaks["bf16"], traffic=traffic, peak_source=peaks["source"] + ", burst (stage timed alone)",
                    share_of_event=d["ms_per_event"] / total,
                    event_tflops=FLOPS_PER_STATE_STEP_PX * hw * hw * B / (total * 1e-3) / 1e12)

    cpu = gpu_eager = module_level = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(hw)
        gpu_eager = gpu_eager_baseline(hw, dev)
    if rank == 0 and world == 1 and args.grid == "cell" and not args.no_module_level:
        module_level = module_level_numbers(model, dev, B)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=warm,
                    ms_per_step=ms_total / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                    data="synthetic", config=workload_config(args, hw, B), clocks=clocks.summary(), e2e=e2e, gpu_launches=launches,
                    roofline=roof, cpu_baseline=cpu, gpu_eager_baseline=gpu_eager, module_level=module_level, stages=stages,
                    events_per_sec=world * (ro.n_state_steps + ro.n_jumps) * args.steps / (ms_total * 1e-3),
                    tflops=world * (ro.n_cell_evals * 2 * (227 * 4096 + 128) + ro.n_prior_evals * 2 * 137 * 4096) * hw * hw * args.steps
                    / (ms_total * 1e-3) / 1e12)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def module_level_numbers(model, dev, B, reps=3):
    """The reference-faithful grid: FuturePredictionODE.forward on a 200x200x64 BEV (-> 50x50x64 latent), same
    observation / target schedule, B samples.  ode_loop = the CUDA rollout alone on the 50x50 latents; forward = the whole
    module call (torch encoder / decoder / refinement around it) with HOST buffers (pinned H2D of the BEV states, D2H of x)."""
    import torch

    H = 200
    g = torch.Generator().manual_seed(3)
    cam_h = torch.randn(B, 3, 64, H, H, generator=g).pin_memory()
    lid_h = torch.randn(B, 5, 64, H, H, generator=g).pin_memory()
    ct = torch.tensor([CAM_T] * B, dtype=torch.float64)
    lt = torch.tensor([LIDAR_T] * B, dtype=torch.float64)
    tt = torch.tensor([TARGETS] * B, dtype=torch.float64)
    fpi = torch.zeros(B, 1, 64, H, H, device=dev)
    out_h = torch.empty((B, len(TARGETS), 64, H, H), dtype=torch.float32).pin_memory()
    ode = model.gru_ode
    times = sorted(CAM_T + LIDAR_T)

    def forward_host():
        with torch.no_grad():
            x, _ = model(fpi, cam_h.to(dev, non_blocking=True), lid_h.to(dev, non_blocking=True), ct, lt, tt)
            out_h.copy_(x, non_blocking=True)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    ms_fwd = timed(forward_host)
    n_steps = ode.last_rollout.n_state_steps
    with torch.no_grad():
        hx = torch.tanh(torch.randn(B * len(times), 64, H // 4, H // 4, device=dev))
        was = ode.cuda_graph
        ode.cuda_graph = False
        ms_eager = timed(lambda: ode.integrate_latents(hx, [len(times)] * B, [times] * B, [TARGETS] * B, 0.05))
        ode.cuda_graph = True
        ms_ode = timed(lambda: ode.integrate_latents(hx, [len(times)] * B, [times] * B, [TARGETS] * B, 0.05))
        ode.cuda_graph = was
    return dict(grid="50x50x64 latent of a 200x200x64 BEV", batch=B, ode_loop_ms=ms_ode, ode_loop_value=n_steps / (ms_ode * 1e-3),
                ode_loop_eager_launch_ms=ms_eager,
                forward_host_buffers_ms=ms_fwd, forward_value=n_steps / (ms_fwd * 1e-3), unit=UNIT,
                note="forward = SmallEncoder -> ODE loop -> SmallDecoder -> SpatialGRU/Block/SpatialGRU/DeepLabHead, every stage on the "
                     "CUDA engine (conv-stage kernels); H2D of the 8 BEV frames / sample and D2H of the 7 output frames inside the time")


def time_stages(eng, B, hw, peaks, reps=10):
    """CUDA-event time of each conv stage's launches, alone, on the launching stream (derivative-cell weights)."""
    import torch
    from streamingflow_b200 import _lib as L
    from streamingflow_b200 import engine as en

    n = B
    evd = dict(kind=0, samples=list(range(n)), x_img=list(range(n)), rec=[-1] * n, eps=[0] * n, dt=[0.1] * n, x_buf=en.BUF_X, s_in=0,
               s_base=0, s_out=0, run_cell=1, run_prior=1, want_f32=0)
    table, evs = eng.build_table([evd])
    tdev = eng.upload_table(table)
    names = dict(eng.stage_names)          # derivative-cell stages, prior-network stages, the two SE layers
    out = {}
    for slot, name in names.items():
        for _ in range(2):
            eng.run_stage(slot, evs[0], tdev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            eng.run_stage(slot, evs[0], tdev)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        if name.startswith("se"):
            # squeeze-excite layer, 128 channels bf16.  Folded (default): ONE streaming pass, the channel-sum reduce reads z once
            # (256 B / pixel), then the tiny scale + weight-fold kernels.  Unfolded: reduce + apply read z twice and write y.
            bytes_px = 256.0 if eng.se_fold else 768.0
            gbs = bytes_px * hw * hw * n / (ms * 1e-3) / 1e9
            out[name] = dict(ms_per_event=ms, gbs=gbs, frac_of_hbm_peak=gbs / peaks["hbm"], bytes_per_pixel=bytes_px)
            continue
        flops = 2.0 * STAGE_MACS[name] * hw * hw * n
        out[name] = dict(ms_per_event=ms, tflops=flops / (ms * 1e-3) / 1e12, frac_of_bf16_peak=flops / (ms * 1e-3) / 1e12 / peaks["bf16"])
    return out


def cpu_baseline(hw):
    """The oracle port (same ATen calls as the reference's modules) on the host cores: ONE sample of the batch, its full
    schedule, repeated until ~10 s of CPU work have been timed."""
    import torch
    from oracle import sf_oracle as so

    torch.set_num_threads(os.cpu_count() or 1)
    m = make_model("cpu")
    sd = {"g." + k: v for k, v in m.gru_ode.state_dict().items()}
    times = sorted(CAM_T + LIDAR_T)
    full = so.build_schedule(times, TARGETS, 0.05, True)
    n_steps = sum(1 for e in full.events if e.kind == "step")
    g = torch.Generator().manual_seed(1)
    hx = torch.tanh(torch.randn(len(times), 64, hw, hw, generator=g))
    eps = (torch.randn(1, 64, hw, hw, generator=g) for _ in range(10 ** 6))
    reps, dt = 0, 0.0
    with torch.no_grad():
        so.integrate_latent(sd, "g", hx, so.Schedule(events=full.events[:2]), eps)      # warm-up
        while dt < 10.0 and reps < 50:
            t0 = time.perf_counter()
            so.integrate_latent(sd, "g", hx, full, eps)
            dt += time.perf_counter() - t0
            reps += 1
    return dict(value=n_steps * reps / dt, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"oracle port, fp32, 1 sample x full schedule ({len(full.events)} events, {n_steps} state-steps) at {hw}x{hw}x64, "
                       f"{reps} repetitions = {dt:.1f} s of CPU work")


def gpu_eager_baseline(hw, dev):
    """The bar on the same B200: the reference algorithm as PyTorch-eager ATen calls (cuDNN, TF32 allowed = torch's
    default), one sample at a time like the reference's loop, timed with CUDA events.  Reported, not part of `value`."""
    import torch
    from oracle import sf_oracle as so

    m = make_model(dev)
    sd = {"g." + k: v for k, v in m.gru_ode.state_dict().items()}
    times = sorted(CAM_T + LIDAR_T)
    full = so.build_schedule(times, TARGETS, 0.05, True)
    n_steps = sum(1 for e in full.events if e.kind == "step")
    hx = torch.tanh(torch.randn(len(times), 64, hw, hw, device=dev))
    eps = (torch.empty(1, 64, hw, hw, device=dev).normal_() for _ in range(10 ** 6))
    with torch.no_grad():
        so.integrate_latent(sd, "g", hx, full, eps)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        reps = 3
        for _ in range(reps):
            so.integrate_latent(sd, "g", hx, full, eps)
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    return dict(value=n_steps / (ms * 1e-3), unit=UNIT, ms_per_sample_rollout=ms, kind="oracle port on cuda (PyTorch eager, cuDNN, allow_tf32 default)",
                note="no host syncs (the reference adds 2-4 .item() syncs per step)")


if __name__ == "__main__":
    main()
