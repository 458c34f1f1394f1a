#!/usr/bin/env python
"""bench.py -- ODE BEV state-steps/s of the GRU-ODE-Bayes integration path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config2|config3|config4|config5]
                    [--grid cell|module] [--batch B] [--precision bf16|bf16x3] [--quick]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1], "config 2" of SURVEY.md 8d): B = 8 samples per GPU, camera 2 Hz (-1, -0.5, 0) +
LiDAR 5 Hz (-0.8 .. 0) observations, 7 targets (-1 .. 2 s), variable-step Euler, IMPUTE: 8 observation jumps + 10 ODE
state-steps per sample.  state-step := one sample x one ode_step call.  A "step" of the bench = one rollout over the batch.

  value     the literal 200x200x64 ODE state (--grid cell): encoded observations resident in HBM; timed: layout pack, noise
            draw, every stage kernel (one replayed CUDA graph), path gather.  CUDA events, max over ranks.
  e2e       the reference-facing call: FuturePredictionODE.forward (future_prediction_ode.py:32-64) on B samples of a
            200x200x64 BEV with HOST (pinned) buffers -- the H2D copy of the camera / LiDAR states and the D2H read of the output
            are inside the timed region of every step (copy streams, double-buffered, overlapping the previous / next step).
            e2e.latent_level keeps round 1's number (integrate_latents_streamed on the 200x200 state) for continuity.
  roofline  dominant conv stage: algorithmic FLOPs (reference MACs x 2, no credit for folded / padded work) / CUDA-event time
            of that stage's launches on the launching stream, vs the measured bf16 GEMM peak (MEASURED_PEAKS.json).
  parity    outside the timed region: sample 0's selected latents of one timed-shape rollout vs the fp64 oracle on the GPU.
  modes / configs / config5_row_sharded   the accurate (bf16x3) mode, BASELINE configs 3 and 4, and config 5
            (400x400x128, ONE grid row-sharded over all N ranks with NCCL halo exchange: strong scaling) in the same run.
  cpu_baseline / --impl reference   the UNMODIFIED reference (baseline/_ref, staged by __graft_entry__.build(); the oracle
            port if that tree is absent) on the host cores: value = its ode_step / gru_obs / infer_state loop on ONE sample of
            the 200x200x64 state, e2e = its FuturePredictionODE.forward on ONE sample of the batch.
"""
import argparse
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
TARGETS_C3 = [-1.0, -0.5, 0.0] + [0.05 * i for i in range(1, 41)]      # config 3: streaming evaluation, 40 steps of 0.05 s
TARGETS_C4 = [-1.0, -0.5, 0.0] + [0.5 * i for i in range(1, 17)]       # config 4: 8 s horizon, 16 targets of 0.5 s
METRIC = "ode_bev_state_steps_per_sec"
UNIT = "state-steps/s"
# reference MACs per output pixel of each conv stage at C = 64 (cell 227 C^2 + 2C, p_model 137 C^2)
STAGE_MACS = {"gates": 4 * 9 * 128 * 64, "propose": 2 * 9 * 128 * 64, "decode": 9 * 64 * 64, "trunk7": 49 * 128 * 64,
              "trunk1": 64 * 64, "trunk": 49 * 128 * 64 + 64 * 64, "mix": 9 * 64 * 64 + 128 * 64 + 2 * 64, "q1": 9 * 64 * 64,
              "q2": 9 * 64 * 128 + 64 * 128, "q3": 9 * 128 * 128, "q4": 9 * 128 * 128, "q5": 9 * 128 * 128}


def flops_per_state_step_px(C):
    return 2 * (364 * C * C + 2 * C)


def rollout_flops_px(ro, C):
    """Algorithmic FLOPs per pixel of the work a compiled rollout EXECUTES: every op runs the cell (227 C^2 + 2 C MACs); the prior
    net (137 C^2 MACs) runs only where its sample is read (rollout.py: not before a jump, not after the last op)."""
    return ro.n_cell_evals * 2 * (227 * C * C + 2 * C) + ro.n_prior_evals * 2 * 137 * C * C


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), hbm=d["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """NVML clocks / throttle reasons sampled every 10 ms DURING the timed region (nvidia-smi as a fallback)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc, self.stop, self.thread = gpu_index, [], None, threading.Event(), None

    def __enter__(self):
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


def bind_to_gpu_cpus(local_rank):
    """Pins this process to the CPUs NVML reports as local to its GPU BEFORE any pinned host memory is allocated, so the
    staging buffers are first-touched on the GPU's own NUMA node (round 1: every rank allocated with default affinity).
    Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            try:
                numa = pynvml.nvmlDeviceGetNumaNodeId(h)
            except Exception:
                numa = None
            return dict(cpus=f"{allowed[0]}-{allowed[-1]} ({len(allowed)})", numa_node=numa)
    except Exception as e:      # no NVML / no permission: keep the default affinity
        return dict(cpus="default", error=str(e)[:80])
    return dict(cpus="default")


def make_model(device, C=64):
    import torch
    from streamingflow_b200.config import ode_cfg
    from streamingflow_b200.models.future_prediction_ode import FuturePredictionODE

    torch.manual_seed(0)
    m = FuturePredictionODE(C, C, 4, ode_cfg(C)).eval()     # random-init weights, torch default init (seed 0)
    return m.to(device)


def make_decoder(device):
    import torch
    from streamingflow_b200.models.decoder import Decoder

    torch.manual_seed(1)
    gates = dict(perceive_hdmap=False, predict_pedestrian=False, predict_instance=False, predict_future_flow=False, planning=False)
    return Decoder(64, 2, 3, 2, gates).eval().to(device)


def workload_config(args, hw, batch):
    return dict(workload=f"config2: Prediction_LC_ODE_Variable ODE head, batch {batch}/GPU, camera 2Hz + LiDAR 5Hz synthetic BEV "
                         f"observations (8 jumps + 10 variable Euler state-steps per sample), ODE grid {hw}x{hw}x64 "
                         f"({'cell-level: the literal 200x200x64 state' if args.grid == 'cell' else 'module-level latent of a 200x200x64 BEV'})",
                grid=args.grid, batch_per_gpu=batch, precision=args.precision, solver="euler", variable_step=True, impute=True,
                prior_net="evaluated where its sample is read (before each ode_step): 10 of the 18 ops of a sample; the reference also evaluates it "
                          "before a jump and after the last op, where nothing reads the result (GRUObservationCell ignores p) -- outputs are "
                          "bit-identical either way; 'all_prior_evaluated' in this line times the rollout with all 18",
                launch="eager" if getattr(args, "no_graph", False) else "one CUDA graph per rollout (all stage launches); layout pack, one-launch noise draw and gathers eager around it",
                l2="inputs + workspace (>1 GB at 200x200, B=8) exceed the 126 MB L2; no explicit flush" if hw >= 200 else
                   "working set fits L2 (module-level latent): L2 flushed by a 256 MB memset between steps",
                parallelism=f"batch-sharded x{args.gpus}, no collective in the data path")


# ------------------------------------------------------------------------------------------------ the reference arm (CPU)
def _reference_modules():
    """(kind, namespace): the unmodified reference when its files are staged under baseline/_ref (or /root/reference exists in
    the builder container), else None -> the oracle port."""
    from oracle import _refimport as ri

    if ri.reference_available():
        return "reference", ri.import_reference(), ri
    return "port", None, ri


def _ref_cell_rollout(nnfo, hx, times, targets, delta_t):
    """The reference's jump / integrate loop (temporal_ode_bayes.py:539-604) driven through ITS OWN inner API -- ode_step,
    gru_obs, infer_state of the unmodified NNFOwithBayesianJumps -- on already-encoded observations hx [n_obs, C, h, w], i.e.
    on the literal 200x200x64 state (forward itself would first pool the grid by 4).  Variable step."""
    import torch

    state = torch.zeros_like(hx[0:1])
    inp = torch.zeros_like(state)
    cur = min(times)
    n_steps = 0
    for i, t_obs in enumerate(times):
        while cur <= t_obs - delta_t:
            state, inp, cur, _, _ = nnfo.ode_step(state, inp, t_obs - cur, cur)
            n_steps += 1
        state, _ = nnfo.gru_obs(state, inp, hx[i:i + 1])
        inp = nnfo.infer_state(state)[0]
    for t_goal in targets:
        while cur < t_goal:
            state, inp, cur, _, _ = nnfo.ode_step(state, inp, t_goal - cur, cur)
            n_steps += 1
    return state, n_steps


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores, all threads, on OUR arm's
    config / metric / unit; each step = a bounded sample of the workload: ONE sample of the batch (the metric is per sample)."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    hw = 200 if args.grid == "cell" else 50
    kind, ref, ri = _reference_modules()
    times = sorted(CAM_T + LIDAR_T)
    g = torch.Generator().manual_seed(1)
    hx = torch.tanh(torch.randn(len(times), 64, hw, hw, generator=g))
    H = 200
    cam = torch.randn(1, 3, 64, H, H, generator=g)
    lid = torch.randn(1, 5, 64, H, H, generator=g)
    ct, lt, tt = (torch.tensor([v], dtype=torch.float64) for v in (CAM_T, LIDAR_T, TARGETS))
    if kind == "reference":
        torch.manual_seed(0)
        m = ref.FuturePredictionODE(64, 64, 4, ri.make_cfg(64)).eval()

        def cell():
            with torch.no_grad():
                return _ref_cell_rollout(m.gru_ode, hx, times, TARGETS, 0.05)[1]

        def forward():
            with torch.no_grad():
                m(torch.zeros(1, 1, 64, H, H), cam, lid, ct, lt, tt)
    else:
        from oracle import sf_oracle as so

        mm = make_model("cpu")
        sd = {k: v for k, v in mm.state_dict().items()}
        sdg = {"g." + k: v for k, v in mm.gru_ode.state_dict().items()}
        sch = so.build_schedule(times, TARGETS, 0.05, True)

        def cell():
            eps = (torch.randn(1, 64, hw, hw, generator=g) for _ in range(10 ** 6))
            with torch.no_grad():
                so.integrate_latent(sdg, "g", hx, sch, eps)
            return sum(1 for e in sch.events if e.kind == "step")

        def forward():
            eps = (torch.randn(1, 64, H // 4, H // 4, generator=g) for _ in range(10 ** 6))
            with torch.no_grad():
                so.future_prediction_forward(sd, cam, lid, ct, lt, tt, 0.05, eps)

    n_steps = 10
    for _ in range(min(args.warmup, 2)):        # CPU: two warm-ups settle the oneDNN primitive caches
        n_steps = cell()
        forward()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cell()
    dt_cell = time.perf_counter() - t0
    n_fwd = max(2, min(args.steps, 8))          # the full forward is ~3x a cell rollout: bounded to keep the arm within minutes
    t0 = time.perf_counter()
    for _ in range(n_fwd):
        forward()
    dt_fwd = time.perf_counter() - t0
    value = n_steps * args.steps / dt_cell
    e2e_value = n_steps * n_fwd / dt_fwd
    src = "UNMODIFIED reference modules (baseline/_ref)" if kind == "reference" else "oracle port (reference tree not staged)"
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * dt_cell / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=workload_config(args, hw, args.batch),
                cpu_baseline=dict(value=value, unit=UNIT, cores=torch.get_num_threads(), kind=kind,
                                  sample=f"{src}, fp32, {torch.get_num_threads()} threads; value: ONE sample of the batch, its full schedule "
                                         f"(8 jumps + {n_steps} state-steps) on the {hw}x{hw}x64 state through ode_step / gru_obs / infer_state, "
                                         f"{args.steps} repetitions; e2e: FuturePredictionODE.forward on ONE sample of a 200x200x64 BEV, {n_fwd} repetitions"),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0, ms_per_forward=1e3 * dt_fwd / n_fwd,
                         api="FuturePredictionODE.forward (reference), B = 1 sample of the batch, CPU tensors"))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
class Ctx:
    pass


def timed_rollouts(fn, steps, warm, barrier, flush=None):
    """warm untimed calls, then ``steps`` timed ones (CUDA events on the current stream), bracketed by barrier + synchronize."""
    import torch

    for _ in range(warm):
        fn()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        if flush is not None:
            flush.zero_()
        a.record()
        fn()
        b.record()
    barrier()
    return sum(a.elapsed_time(b) for a, b in ev)


def max_over_ranks(ms, dev, world):
    import torch
    import torch.distributed as dist

    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4", "config5"],
                    help="config2 (default) prints the full line incl. the other configs as extra keys; the others print their own line")
    ap.add_argument("--grid", default="cell", choices=["cell", "module"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--quick", action="store_true", help="value / e2e / roofline only (A/B runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stage-timing", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip modes / configs 3-5 / DMA probe")
    ap.add_argument("--no-module-level", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the stages eagerly instead of replaying the captured CUDA graph")
    args = ap.parse_args()
    if args.quick:
        args.no_cpu_baseline = args.no_extras = args.no_module_level = True
    if args.impl == "reference":
        return run_reference_arm(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    affinity = bind_to_gpu_cpus(local)          # before torch allocates any pinned memory

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warm = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    c = Ctx()
    c.args, c.world, c.rank, c.local, c.dev, c.warm, c.barrier = args, world, rank, local, dev, warm, barrier
    if args.workload == "config5":
        line = config5_row_sharded(c, steps=args.steps, full_line=True)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    hw = 200 if args.grid == "cell" else 50
    B = args.batch
    targets = {"config2": TARGETS, "config3": TARGETS_C3, "config4": TARGETS_C4}[args.workload]
    model = make_model(dev)
    ode = model.gru_ode
    ode.precision = args.precision
    ode.cuda_graph = not args.no_graph       # the whole step loop is one captured CUDA graph (eager launches with --no-graph)
    times = sorted(CAM_T + LIDAR_T)
    n_obs = len(times)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    hx_dev = torch.tanh(torch.randn(B * n_obs, 64, hw, hw, device=dev, generator=g))      # encoded observations (tanh head range)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if hw < 200 else None

    def rollout(hx=hx_dev, tg=targets, b=B):
        with torch.no_grad():
            return ode.integrate_latents(hx, [n_obs] * b, [times] * b, [tg] * b, 0.05)

    # ---------------- device-resident timing (value)
    for _ in range(warm):
        rollout()
    eng = ode._engines[next(iter(ode._engines))]["engine"]
    ro = ode.last_rollout
    steps_per_rollout = ro.n_state_steps
    with ClockSampler(local) as clocks:
        ms = timed_rollouts(rollout, args.steps, 0, barrier, flush)
    launches_per_rollout = ro.launches + 4      # stage kernels of one rollout + layout pack + noise fill + path gather + final-state gather
    eng.check_errflag()
    ms_total = max_over_ranks(ms, dev, world)
    value = world * steps_per_rollout * args.steps / (ms_total * 1e-3)

    # the same rollout with the prior net evaluated after EVERY op, as the reference does (the dead evaluations included)
    ode.skip_dead_prior = False
    for _ in range(warm):
        rollout()
    ro_all = ode.last_rollout
    ms_all = max_over_ranks(timed_rollouts(rollout, min(args.steps, 5), 0, barrier, flush), dev, world)
    all_prior = dict(value=world * steps_per_rollout * min(args.steps, 5) / (ms_all * 1e-3), unit=UNIT, ms_per_step=ms_all / min(args.steps, 5),
                     prior_evaluations_per_sample=ro_all.n_prior_evals // B, live_prior_evaluations_per_sample=ro.n_prior_evals // B)
    ode.skip_dead_prior = True

    # ---------------- end to end through FuturePredictionODE.forward with host buffers (e2e)
    e2e = e2e_forward_host(c, model, B, args.steps)
    if args.workload == "config2" and args.grid == "cell" and not args.quick:
        e2e["latent_level"] = e2e_latent_streamed(c, ode, hx_dev, B, n_obs, times, hw, steps_per_rollout, min(args.steps, 5))
        e2e["to_occupancy_masks"] = e2e_forward_host(c, model, B, min(args.steps, 5), decoder=make_decoder(dev))

    # ---------------- per-stage roofline (rank 0)
    peaks = load_peaks()
    roof, stages = None, None
    if rank == 0 and not args.no_stage_timing:
        stages = time_stages(eng, B, hw, peaks)
        roof = roofline_from_stages(stages, peaks, hw, B, 64)

    parity = parity_check(c, ode, hx_dev, B, n_obs, times, targets, hw) if rank == 0 else None

    modes = configs = config5 = dma = None
    if not args.no_extras and args.workload == "config2" and args.grid == "cell":
        modes = dict(bf16x3=accurate_mode(c, model, hx_dev, B, n_obs, times, hw, peaks))
        del hx_dev
        torch.cuda.empty_cache()
        configs = dict(config3=other_config(c, model, "config3", 16, TARGETS_C3), config4=other_config(c, model, "config4", 64, TARGETS_C4))
        ode._engines.clear(); ode._graphs.clear()
        torch.cuda.empty_cache()
        config5 = config5_row_sharded(c, steps=3)
        dma = host_dma_probe(c)

    cpu = gpu_eager = module_level = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(hw)
        gpu_eager = gpu_eager_baseline(hw, dev)
    if rank == 0 and world == 1 and args.grid == "cell" and not args.no_module_level:
        module_level = module_level_numbers(model, dev, B)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=warm,
                    ms_per_step=ms_total / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="bf16" if args.precision == "bf16" else "bf16x3 (split-bf16 operands, 3 products, fp32 accumulate)",
                    data="synthetic", config=workload_config(args, hw, B), clocks=clocks.summary(), e2e=e2e,
                    gpu_launches=launches_per_rollout * args.steps, gpu_launches_per_step=launches_per_rollout,
                    roofline=roof, cpu_baseline=cpu, parity=parity, all_prior_evaluated=all_prior, modes=modes, configs=configs, config5_row_sharded=config5,
                    host_dma_probe=dma, host_affinity=affinity, gpu_eager_baseline=gpu_eager, module_level=module_level, stages=stages,
                    events_per_sec=world * (ro.n_state_steps + ro.n_jumps) * args.steps / (ms_total * 1e-3),
                    tflops=world * (ro.n_cell_evals * 2 * (227 * 4096 + 128) + ro.n_prior_evals * 2 * 137 * 4096) * hw * hw * args.steps
                    / (ms_total * 1e-3) / 1e12)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def e2e_forward_host(c, model, B, steps, decoder=None):
    """The call a user of the reference makes -- FuturePredictionODE.forward(future_prediction_input, camera_states,
    lidar_states, camera_timestamp, lidar_timestamp, target_timestamp) -- with the BEV states in pinned HOST memory and the
    output read back to pinned host memory, every step: H2D on a copy stream into one of two device buffers (the copy of
    step i+1 runs under the compute of step i), forward on the main stream, D2H on a second copy stream."""
    import torch

    dev, H = c.dev, 200
    g = torch.Generator().manual_seed(3 + c.rank)
    cam_h = torch.randn(B, 3, 64, H, H, generator=g).pin_memory()
    lid_h = torch.randn(B, 5, 64, H, H, generator=g).pin_memory()
    out_h = (torch.empty((B, len(TARGETS), 64, H, H), dtype=torch.float32) if decoder is None else
             torch.empty((B, len(TARGETS), H, H), dtype=torch.uint8)).pin_memory()
    ct = torch.tensor([CAM_T] * B, dtype=torch.float64)
    lt = torch.tensor([LIDAR_T] * B, dtype=torch.float64)
    tt = torch.tensor([TARGETS] * B, dtype=torch.float64)
    fpi = torch.zeros(B, 1, 64, H, H, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    bufs = [(torch.empty_like(cam_h, device=dev), torch.empty_like(lid_h, device=dev)) for _ in range(2)]
    marks = []                 # (forward start, forward end) events of every step on the main stream
    up = [None, None]          # H2D-done events per buffer
    free = [None, None]        # forward-has-consumed events per buffer
    out_done = [None]

    def upload(i):
        k = i & 1
        if free[k] is not None:
            s_in.wait_event(free[k])
        with torch.cuda.stream(s_in):
            bufs[k][0].copy_(cam_h, non_blocking=True)
            bufs[k][1].copy_(lid_h, non_blocking=True)
            e = torch.cuda.Event()
            e.record(s_in)
        up[k] = e

    def step(i, last):
        k = i & 1
        if not last:
            upload(i + 1)
        main.wait_event(up[k])
        t0 = torch.cuda.Event(enable_timing=True)
        t0.record(main)
        with torch.no_grad():
            x, _ = model(fpi, bufs[k][0], bufs[k][1], ct, lt, tt)
            if decoder is not None:      # the step after the head: BEV Decoder + arg-max on the engine, frames handed over in engine layout
                x = decoder(x, planes=model.last_output_planes)["segmentation_argmax"].view(B, len(TARGETS), H, H)
        e = torch.cuda.Event(enable_timing=True)
        e.record(main)
        free[k] = e
        marks.append((t0, e))
        if out_done[0] is not None:
            s_out.wait_event(out_done[0])
        s_out.wait_event(e)
        with torch.cuda.stream(s_out):
            out_h.copy_(x, non_blocking=True)
            x.record_stream(s_out)
            d = torch.cuda.Event()
            d.record(s_out)
        out_done[0] = d

    def run(n):
        upload(0)
        for i in range(n):
            step(i, i == n - 1)
        main.wait_event(out_done[0])

    run(2)      # warm-up: builds the codec / engine / refinement plans
    c.barrier()
    marks.clear()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(steps)
    b.record()
    c.barrier()
    ms = max_over_ranks(a.elapsed_time(b), dev, c.world)
    # where a step's time goes on the main stream: the forward itself, and the wait for the step's H2D copy in front of it
    fwd_ms = statistics.mean(s0.elapsed_time(s1) for s0, s1 in marks)
    stall_ms = statistics.mean(marks[i][1].elapsed_time(marks[i + 1][0]) for i in range(len(marks) - 1)) if len(marks) > 1 else 0.0
    n_steps = model.gru_ode.last_rollout.n_state_steps
    if decoder is not None:
        return dict(value=c.world * n_steps * steps / (ms * 1e-3), unit=UNIT, h2d_bytes_per_step=(cam_h.numel() + lid_h.numel()) * 4,
                    d2h_bytes_per_step=out_h.numel(), ms_per_step=ms / steps, steps=steps, forward_ms_in_pipeline=fwd_ms, wait_for_h2d_ms=stall_ms,
                    api="FuturePredictionODE.forward -> Decoder.forward(x, planes=...) -> segmentation arg-max (trainer.py:230-231), all on the "
                        "engine; H2D of the BEV states and D2H of the uint8 occupancy masks [B, 7, 200, 200] inside the timed region")
    return dict(value=c.world * n_steps * steps / (ms * 1e-3), unit=UNIT, h2d_bytes_per_step=(cam_h.numel() + lid_h.numel()) * 4,
                d2h_bytes_per_step=out_h.numel() * 4, ms_per_step=ms / steps, steps=steps,
                launches_per_step=model.gru_ode.last_rollout.launches, forward_ms_in_pipeline=fwd_ms, wait_for_h2d_ms=stall_ms,
                api="FuturePredictionODE.forward(future_prediction_input, camera_states, lidar_states, camera_timestamp, lidar_timestamp, "
                    f"target_timestamp), B = {B}/GPU, BEV 200x200x64 (-> 50x50x64 latent), camera/LiDAR states in pinned host memory, output "
                    "read back to pinned host memory; all copies of every timed step inside the timed region (double-buffered copy streams)")


def e2e_latent_streamed(c, ode, hx_dev, B, n_obs, times, hw, steps_per_rollout, n):
    """Round 1's e2e, kept as an extra key: the 200x200x64 state through integrate_latents_streamed with pinned host buffers."""
    import torch

    hx_host = hx_dev.cpu().pin_memory()
    sel_host = torch.empty((len(TARGETS), B, 64, hw, hw), dtype=torch.float32).pin_memory()

    def once():
        with torch.no_grad():
            ode.integrate_latents_streamed(hx_host, [n_obs] * B, [times] * B, [TARGETS] * B, 0.05, out_host=sel_host, join=False)

    once()
    c.barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for _ in range(n):
        once()
    torch.cuda.current_stream().wait_event(ode.download_done)
    eb.record()
    c.barrier()
    ms = max_over_ranks(ea.elapsed_time(eb), c.dev, c.world)
    return dict(value=c.world * steps_per_rollout * n / (ms * 1e-3), unit=UNIT, steps=n, h2d_bytes_per_step=hx_host.numel() * 4,
                d2h_bytes_per_step=sel_host.numel() * 4,
                api="NNFOwithBayesianJumps.integrate_latents_streamed on the 200x200x64 state, pinned host buffers (not a reference API)")


def roofline_from_stages(stages, peaks, hw, B, C):
    total = sum(s["ms_per_event"] for s in stages.values())
    dom = max((k for k in stages if "tflops" in stages[k]), key=lambda k: stages[k]["ms_per_event"])
    d = stages[dom]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"{dom}@{hw}x{B}")
    return dict(bound="tensor", kernel=f"conv_stage_kernel<{dom}>", achieved=d["tflops"], peak=peaks["bf16"], unit="TFLOP/s",
                frac=d["tflops"] / peaks["bf16"], traffic=traffic, peak_source=peaks["source"] + ", burst (stage timed alone)",
                share_of_event=d["ms_per_event"] / total,
                event_tflops=flops_per_state_step_px(C) * hw * hw * B / (total * 1e-3) / 1e12,
                event_frac_of_burst=flops_per_state_step_px(C) * hw * hw * B / (total * 1e-3) / 1e12 / peaks["bf16"],
                event_frac_of_sustained=flops_per_state_step_px(C) * hw * hw * B / (total * 1e-3) / 1e12 / peaks["bf16_sustained"])


def parity_check(c, ode, hx_dev, B, n_obs, times, targets, hw):
    """Outside the timed region: one rollout of the TIMED shape (same weights, same observations, graph replay) on a known
    noise tape; sample 0's selected latents against the fp64 oracle evaluated on the GPU (oracle = checker only)."""
    import torch
    from oracle import sf_oracle as so

    try:
        sch = so.build_schedule(times, targets, 0.05, True)
        n_eps = len(sch.events)
        gen = torch.Generator(device=c.dev).manual_seed(777)
        tape = torch.randn(n_eps * B, 64, hw, hw, device=c.dev, generator=gen)
        orig = ode.__dict__.get("_draw_noise")
        ode._draw_noise = lambda n, h, w, device, out=None: tape[:max(n, 1)]
        try:
            with torch.no_grad():
                _, sel = ode.integrate_latents(hx_dev, [n_obs] * B, [times] * B, [targets] * B, 0.05)
        finally:
            if orig is None:
                ode.__dict__.pop("_draw_noise", None)
            else:
                ode._draw_noise = orig
        sd = {"g." + k: (v.double() if v.is_floating_point() else v) for k, v in ode.state_dict().items()}
        a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            with torch.no_grad():
                _, path = so.integrate_latent(sd, "g", hx_dev[:n_obs].double(), sch, iter(tape[:n_eps].double()[:, None]))
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b
        want = torch.stack([path[i][0] for i in sch.select])
        err = ((sel[0].double() - want).abs().max() / want.abs().max()).item()
        tol = 1e-2 if ode.precision == "bf16" else 1e-4
        return dict(checked="sample 0 of the timed batch: the 7 selected latent states of the full rollout (18 events) vs the fp64 oracle",
                    max_rel_err=err, tolerance=tol, ok=bool(err < tol), precision=ode.precision)
    except Exception as e:      # the bench line must still be printed
        return dict(error=str(e)[:200])


def accurate_mode(c, model, hx_dev, B, n_obs, times, hw, peaks):
    """The 1e-4 path (split-bf16 operands, three products) on the same workload: throughput and the dominant stage's roofline."""
    import torch

    ode = model.gru_ode
    was = ode.precision
    ode.precision = "bf16x3"
    try:
        def rollout():
            with torch.no_grad():
                return ode.integrate_latents(hx_dev, [n_obs] * B, [times] * B, [TARGETS] * B, 0.05)

        steps = 3
        ms = max_over_ranks(timed_rollouts(rollout, steps, 2, c.barrier), c.dev, c.world)
        n = ode.last_rollout.n_state_steps
        out = dict(value=c.world * n * steps / (ms * 1e-3), unit=UNIT, ms_per_step=ms / steps, steps=steps,
                   dtype="bf16x3: x*w ~ xh*wh + xh*wl + xl*wh, fp32 accumulate (kind::tf32 operands measure 5e-4 on this rollout, outside 1e-4)",
                   tflops_algorithmic=c.world * rollout_flops_px(ode.last_rollout, 64) * hw * hw * steps / (ms * 1e-3) / 1e12)
        if c.rank == 0 and not c.args.no_stage_timing:
            eng = ode._engines[(str(c.dev), hw, hw, "bf16x3")]["engine"]
            st = time_stages(eng, B, hw, peaks, reps=5)
            out["roofline"] = roofline_from_stages(st, peaks, hw, B, 64)
            out["roofline"]["note"] = "algorithmic FLOPs (one product per MAC); the tensor pipe executes three"
            out["parity"] = parity_check(c, ode, hx_dev, B, n_obs, times, TARGETS, hw)
        ode._engines.pop((str(c.dev), hw, hw, "bf16x3"), None)
        return out
    finally:
        ode.precision = was


def other_config(c, model, name, total_batch, targets):
    """BASELINE configs 3 / 4 at the cell level (200x200x64 state): the TOTAL batch is split over the ranks (16 / N, 64 / N)."""
    import torch

    ode = model.gru_ode
    b = max(1, total_batch // c.world)
    times = sorted(CAM_T + LIDAR_T)
    n_obs, hw = len(times), 200
    ode._engines.clear(); ode._graphs.clear()
    torch.cuda.empty_cache()
    g = torch.Generator(device=c.dev).manual_seed(11 + c.rank)
    hx = torch.tanh(torch.randn(b * n_obs, 64, hw, hw, device=c.dev, generator=g))

    def rollout():
        with torch.no_grad():
            return ode.integrate_latents(hx, [n_obs] * b, [times] * b, [targets] * b, 0.05)

    steps = 2
    ms = max_over_ranks(timed_rollouts(rollout, steps, 2, c.barrier), c.dev, c.world)
    ro = ode.last_rollout
    return dict(workload=f"{name}: batch {total_batch} total = {b}/GPU x {c.world} GPU, {len(targets)} targets, "
                         f"{ro.n_state_steps // b} state-steps + {ro.n_jumps // b} jumps per sample, 200x200x64 state, bf16",
                value=c.world * ro.n_state_steps * steps / (ms * 1e-3), unit=UNIT, ms_per_step=ms / steps, steps=steps, scaling="strong",
                tflops_algorithmic=c.world * rollout_flops_px(ro, 64) * hw * hw * steps / (ms * 1e-3) / 1e12)


def config5_row_sharded(c, steps=3, full_line=False):
    """BASELINE config 5: ONE 400x400x128 grid (B = 1), its rows split over all N ranks with a 12-row halo exchange (NCCL
    send/recv) and two [B, 2C] all-reduces per event (row_sharding.py).  Strong scaling: the same work at every N."""
    import torch
    from oracle import sf_oracle as so
    from streamingflow_b200.config import ode_cfg
    from streamingflow_b200.layers.temporal_ode_bayes import NNFOwithBayesianJumps
    from streamingflow_b200.row_sharding import RowShardedOde

    C, H = 128, 400
    torch.manual_seed(0)
    m = NNFOwithBayesianJumps(C, C, ode_cfg(C)).eval().to(c.dev)
    m.precision = "bf16"
    times = sorted(CAM_T + LIDAR_T)
    g = torch.Generator(device=c.dev).manual_seed(5)        # the same full grid on every rank
    hx = torch.tanh(torch.randn(len(times), C, H, H, device=c.dev, generator=g))
    sh = RowShardedOde(m, H, H, 1)
    n_eps = 18
    tape = torch.randn(n_eps, C, H, H, device=c.dev, generator=g)
    # every rank keeps only its local image (band + halo rows) of the observations and of the noise tape resident, the way a
    # sharded producer delivers them; the values are those of the same full grid on every rank
    hx, tape = hx[:, :, sh.lo:sh.hi].contiguous(), tape[:, :, sh.lo:sh.hi].contiguous()

    def rollout():
        with torch.no_grad():
            return sh.integrate(hx, [len(times)], [times], [TARGETS], 0.05, noise=tape, local_rows=True)

    rollout(); rollout()
    c.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(c.local) as clocks:
        a.record()
        for _ in range(steps):
            band, ro = rollout()
        b.record()
        c.barrier()
    ms = max_over_ranks(a.elapsed_time(b), c.dev, c.world)
    n = ro.n_state_steps
    value = n * steps / (ms * 1e-3)
    flops = rollout_flops_px(ro, C) * H * H
    out = dict(workload=f"config5: 400x400x128 ODE state, B = 1, row-sharded over {c.world} GPU(s) ({sh.own_hi - sh.own_lo} rows + 12-row halos per rank), "
                        "per event one halo push / pull per neighbour + two [B,2C] all-reduces (NVLink peer-memory kernels, or NCCL calls: see transport), "
                        "replayed as CUDA graph(s)",
               value=value, unit=UNIT, ms_per_rollout=ms / steps, steps=steps, scaling="strong", n_gpus=c.world,
               tflops_algorithmic=flops * steps / (ms * 1e-3) / 1e12, band_rows=sh.own_hi - sh.own_lo, halo_rows=12, launch=sh.graph_mode,
               transport=sh.transport, peer_error=sh.peer_error, whole_graph_error=sh.__dict__.get("_whole_graph_error"))
    launches = sh.launches
    if os.environ.get("SF_ROWSHARD_PROFILE", "0") == "1":      # one more rollout with CUDA-event marks between its phases
        sh.profile = []
        t0 = time.perf_counter()
        rollout()
        host_ms = (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        marks = sh.profile
        sh.profile = None
        out["phases_ms"] = {f"{a[0]} -> {b[0]}": round(a[1].elapsed_time(b[1]), 3) for a, b in zip(marks[:-1], marks[1:])}
        out["phases_ms"]["host time of the call"] = round(host_ms, 3)
    if sh.arena is not None and sh.arena.tracing:          # SF_PEER_TRACE=1: where the exchange time goes (in-kernel %globaltimer stamps)
        out["peer_trace"] = sh.arena.trace_report()
        if c.rank == c.world // 2 and c.rank != 0:
            print(json.dumps(dict(rank=c.rank, peer_trace=out["peer_trace"])), file=sys.stderr, flush=True)
    sh.release_graphs()          # graphs with captured NCCL kernels must be destroyed before the process group
    del sh
    if not full_line:
        return out
    peaks = load_peaks()
    return dict(metric=METRIC, value=value, unit=UNIT, n_gpus=c.world, steps=steps, warmup=2, ms_per_step=ms / steps, higher_is_better=True,
                scaling="strong", vs_baseline=None, dtype="bf16", data="synthetic", config=dict(workload=out["workload"]), clocks=clocks.summary(),
                gpu_launches=launches, e2e=None, cpu_baseline=None,
                roofline=dict(bound="tensor", kernel="whole event (all conv stages)", achieved=out["tflops_algorithmic"] / c.world, peak=peaks["bf16"],
                              unit="TFLOP/s", frac=out["tflops_algorithmic"] / c.world / peaks["bf16"], traffic=None,
                              peak_source=peaks["source"] + "; per GPU, algorithmic FLOPs of the owned rows only"))


def host_dma_probe(c, mb=256, reps=6):
    """All ranks at once: pinned host -> device and device -> pinned host copies of ``mb`` MB on two streams, concurrently in
    both directions (what e2e does per step).  The aggregate over ranks is the host-side DMA ceiling of this box."""
    import torch

    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=c.dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=c.dev)
    s1, s2 = torch.cuda.Stream(c.dev), torch.cuda.Stream(c.dev)
    res = {}
    for mode in ("h2d", "d2h", "both"):
        c.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_event(a); s2.wait_event(a)
        for _ in range(reps):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        b.record()
        c.barrier()
        ms = max_over_ranks(a.elapsed_time(b), c.dev, c.world)
        per_gpu = (2 if mode == "both" else 1) * n * reps / (ms * 1e-3) / 1e9
        res[mode] = dict(gbs_per_gpu=per_gpu, gbs_aggregate=per_gpu * c.world)
    res["note"] = f"{mb} MB pinned buffers, {reps} copies per direction, all {c.world} rank(s) concurrently; 'both' = the sum of the two directions"
    return res


def module_level_numbers(model, dev, B, reps=3):
    """The reference-faithful grid: the ODE loop alone on the 50x50x64 latents of a 200x200x64 BEV (graph replay vs
    eager launches).  The whole forward with host buffers is the line's e2e."""
    import torch

    H = 200
    ode = model.gru_ode
    times = sorted(CAM_T + LIDAR_T)

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

    with torch.no_grad():
        hx = torch.tanh(torch.randn(B * len(times), 64, H // 4, H // 4, device=dev))
        was = ode.cuda_graph
        ode.cuda_graph = False
        ms_eager = timed(lambda: ode.integrate_latents(hx, [len(times)] * B, [times] * B, [TARGETS] * B, 0.05))
        ode.cuda_graph = True
        ms_ode = timed(lambda: ode.integrate_latents(hx, [len(times)] * B, [times] * B, [TARGETS] * B, 0.05))
        ode.cuda_graph = was
    n_steps = ode.last_rollout.n_state_steps
    return dict(grid="50x50x64 latent of a 200x200x64 BEV", batch=B, ode_loop_ms=ms_ode, ode_loop_value=n_steps / (ms_ode * 1e-3),
                ode_loop_eager_launch_ms=ms_eager, unit=UNIT)


def time_stages(eng, B, hw, peaks, reps=10):
    """CUDA-event time of each conv stage's launches, alone, on the launching stream (derivative-cell weights)."""
    import torch
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
    """The reference on the host cores (unmodified modules from baseline/_ref; the oracle port if absent), bounded samples:
    (i) ONE sample of the batch, its full schedule on the hw x hw x 64 state, repeated until ~10 s of CPU work have been timed;
    (ii) BASELINE config 1: the full FuturePredictionODE.forward, B = 1, 3 camera frames, 4 future targets (median of 5 after 2
    warm-ups); (iii) one ode_step at 50x50 and 200x200."""
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    kind, ref, ri = _reference_modules()
    times = sorted(CAM_T + LIDAR_T)
    g = torch.Generator().manual_seed(1)
    hx = torch.tanh(torch.randn(len(times), 64, hw, hw, generator=g))
    extra = {}
    if kind == "reference":
        torch.manual_seed(0)
        m = ref.FuturePredictionODE(64, 64, 4, ri.make_cfg(64)).eval()
        reps, dt, n_steps = 0, 0.0, 10
        with torch.no_grad():
            m.gru_ode.ode_step(torch.zeros(1, 64, hw, hw), torch.zeros(1, 64, hw, hw), 0.1, 0.0)      # warm-up
            while dt < 10.0 and reps < 50:
                t0 = time.perf_counter()
                n_steps = _ref_cell_rollout(m.gru_ode, hx, times, TARGETS, 0.05)[1]
                dt += time.perf_counter() - t0
                reps += 1
            H = 200
            cam = torch.randn(1, 3, 64, H, H, generator=g)
            ct = torch.tensor([CAM_T], dtype=torch.float64)
            tt = torch.tensor([[0.5, 1.0, 1.5, 2.0]], dtype=torch.float64)
            ts = []
            for i in range(7):
                t0 = time.perf_counter()
                m(torch.zeros(1, 1, 64, H, H), cam, None, ct, None, tt)
                ts.append(time.perf_counter() - t0)
            extra["config1_forward_s"] = statistics.median(ts[2:])
            extra["config1"] = "FuturePredictionODE.forward, B=1, 200x200x64 BEV, 3 camera frames, 4 x 0.5 s targets (3 jumps + 6 state-steps), median of 5 after 2 warm-ups"
            for s in (50, 200):
                st, inp = torch.zeros(1, 64, s, s), torch.zeros(1, 64, s, s)
                m.gru_ode.ode_step(st, inp, 0.1, 0.0)
                t0 = time.perf_counter()
                k = 10 if s == 50 else 3
                for _ in range(k):
                    m.gru_ode.ode_step(st, inp, 0.1, 0.0)
                extra[f"ode_step_ms_{s}x{s}"] = 1e3 * (time.perf_counter() - t0) / k
        n_events = 8 + n_steps
    else:
        from oracle import sf_oracle as so

        mm = make_model("cpu")
        sd = {"g." + k: v for k, v in mm.gru_ode.state_dict().items()}
        full = so.build_schedule(times, TARGETS, 0.05, True)
        n_steps = sum(1 for e in full.events if e.kind == "step")
        n_events = len(full.events)
        eps = (torch.randn(1, 64, hw, hw, generator=g) for _ in range(10 ** 6))
        reps, dt = 0, 0.0
        with torch.no_grad():
            so.integrate_latent(sd, "g", hx, so.Schedule(events=full.events[:2]), eps)      # warm-up
            while dt < 10.0 and reps < 50:
                t0 = time.perf_counter()
                so.integrate_latent(sd, "g", hx, full, eps)
                dt += time.perf_counter() - t0
                reps += 1
    return dict(value=n_steps * reps / dt, unit=UNIT, cores=torch.get_num_threads(), kind=kind,
                sample=f"{'unmodified reference (baseline/_ref)' if kind == 'reference' else 'oracle port'}, fp32, 1 sample x full schedule "
                       f"({n_events} events, {n_steps} state-steps) at {hw}x{hw}x64, {reps} repetitions = {dt:.1f} s of CPU work",
                parallel_info=torch.__config__.parallel_info().split("\n")[0], **extra)


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
