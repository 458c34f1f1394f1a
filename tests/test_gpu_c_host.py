"""GPU parity of the ODE head driven through include/sf_b200.h ALONE (pytest -m gpu): sf_ode_query_workspace / sf_ode_create
(library-side weight packing and workspace carving) / sf_rollout_plan_create (library-side schedule) / sf_ode_rollout /
sf_ode_read_path -- no OdeEngine, no Python packing, no Python schedule -- against the fp64 oracle and, bit for bit, against the
Python module's engine path.  tests/c_host/rollout_host.c is the same sequence as a C program (built with gcc against the header)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import sf_oracle as so
from oracle._refimport import make_cfg
from streamingflow_b200 import _lib as L
from streamingflow_b200 import cpack

pytestmark = pytest.mark.gpu
TOL = {"bf16": 1e-2, "bf16x3": 1e-4}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def _case(Cc, h, w, B, seed, jitter=True):
    rng = np.random.RandomState(seed)
    base = sorted([-1.0, -0.5, 0.0] + [-0.8, -0.6, -0.4, -0.2, 0.0])
    tgt = [-1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0]
    j = (lambda: float(rng.uniform(-0.02, 0.02))) if jitter else (lambda: 0.0)
    times = [sorted(t + j() for t in base) for _ in range(B)]
    targets = [sorted(t + j() for t in tgt) for _ in range(B)]
    g = torch.Generator().manual_seed(seed)
    hx = torch.tanh(torch.randn(B * len(base), Cc, h, w, generator=g))
    return times, targets, hx


class COde:
    """The header's call sequence through ctypes: nothing but libsf_b200.so entry points and raw device pointers."""

    def __init__(self, sd, Cc, h, w, B, precision, n_path, n_obs, n_eps, dev):
        self.lib = lib = L.load()
        self.geo = L.Geometry(B, h, w, Cc, L.PREC_BF16X3 if precision == "bf16x3" else L.PREC_BF16, dev.index or 0)
        self.opt = L.OdeOptions(n_path, n_obs, n_eps, L.PACK_DEFAULT)     # the engine's defaults
        nbytes = C.c_size_t()
        L.check(lib.sf_ode_query_workspace(C.byref(self.geo), C.byref(self.opt), C.byref(nbytes)), "sf_ode_query_workspace")
        self.ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)          # the caller owns the device memory
        arr, n, keep = cpack.tensor_table(sd)
        self.h = C.c_void_p()
        L.check(lib.sf_ode_create(C.byref(self.geo), C.byref(self.opt), arr, n, b"", self.ws.data_ptr(), nbytes.value, C.byref(self.h)), "sf_ode_create")
        self.dev, self.dims = dev, (Cc, h, w)

    def tensor(self, which, dtype):
        p, n = C.c_void_p(), C.c_size_t()
        L.check(self.lib.sf_ode_tensor(self.h, which, C.byref(p), C.byref(n)), "sf_ode_tensor")
        off = p.value - self.ws.data_ptr()
        return self.ws[off:off + n.value].view(dtype)

    def run(self, hx, tape, plan):
        lib, stream = self.lib, C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        Cc, h, w = self.dims
        L.check(lib.sf_ode_set_observations(self.h, hx.data_ptr(), 0, hx.shape[0], stream), "sf_ode_set_observations")
        self.tensor(L.ODE_EPS, torch.float32)[:tape.numel()].copy_(tape.reshape(-1))
        L.check(lib.sf_ode_reset_state(self.h, stream), "sf_ode_reset_state")
        table = torch.from_numpy(plan.table).to(self.dev)
        evs = (L.Event * len(plan.events))(*plan.events)
        L.check(lib.sf_ode_rollout(self.h, evs, len(plan.events), table.data_ptr(), stream), "sf_ode_rollout")
        flat = torch.tensor([s for row in plan.out_slots for s in row], dtype=torch.int32, device=self.dev)
        out = torch.empty((flat.numel(), Cc, h, w), dtype=torch.float32, device=self.dev)
        L.check(lib.sf_ode_read_path(self.h, flat.data_ptr(), flat.numel(), out.data_ptr(), stream), "sf_ode_read_path")
        torch.cuda.synchronize()
        assert int(self.tensor(L.ODE_ERRFLAG, torch.int32)[0].item()) == 0
        return out

    def close(self):
        L.check(self.lib.sf_ode_destroy(self.h), "sf_ode_destroy")


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
@pytest.mark.parametrize("Cc,solver", [(64, "euler"), (64, "midpoint"), (128, "euler")])
def test_ode_head_from_the_header_alone_matches_oracle_and_engine(precision, Cc, solver):
    from streamingflow_b200.layers.temporal_ode_bayes import NNFOwithBayesianJumps

    dev = torch.device("cuda", 0)
    h, w, B, seed = 20, 13, 3, 11
    times, targets, hx = _case(Cc, h, w, B, seed)
    m = NNFOwithBayesianJumps(Cc, Cc, make_cfg(Cc, solver=solver)).eval()
    sd = so.recipe_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed, 1.0)
    m.load_state_dict(sd, strict=True)
    plan = cpack.plan_rollout(times, targets, 0.05, True, solver)                 # the library's own schedule
    tape = torch.randn(plan.info["n_eps"], Cc, h, w, generator=torch.Generator().manual_seed(seed + 1))
    hot = {k: v for k, v in sd.items() if k.startswith(("gru_c.", "gru_obs.", "p_model."))}
    ode = COde(hot, Cc, h, w, B, precision, plan.info["n_path"], hx.shape[0], plan.info["n_eps"], dev)
    got = ode.run(hx.to(dev), tape.to(dev), plan).view(B, len(targets[0]), Cc, h, w)
    ode.close()
    # (1) the fp64 oracle, sample by sample (noise slots are sample-major)
    sd64 = {"g." + k: (v.double().to(dev) if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
    n_obs, off = len(times[0]), 0
    with torch.no_grad():
        for b in range(B):
            sch = so.build_schedule(times[b], targets[b], 0.05, True)
            per = 2 if solver == "midpoint" else 1
            n_noise = sum(per if e.kind == "step" else 1 for e in sch.events)
            eps = iter(tape[off:off + n_noise, None].double().to(dev))
            off += n_noise
            _, path = so.integrate_latent(sd64, "g", hx[b * n_obs:(b + 1) * n_obs].double().to(dev), sch, eps, solver)
            ref = torch.cat([path[i] for i in sch.select])
            err = _rel(got[b], ref)
            assert err < TOL[precision], f"sample {b}: selected latents vs fp64 oracle {err:.3e} ({precision}, C={Cc}, {solver})"
    assert off == plan.info["n_eps"]
    # (2) the Python module (schedule.py, rollout.py, OdeEngine) on the same inputs: bit-identical
    m = m.to(dev)
    m.precision = precision
    m.cuda_graph = False
    m._draw_noise = lambda n, hh, ww, device: tape[:max(n, 1)].to(device).contiguous()
    with torch.no_grad():
        _, sel = m._integrate_impl(hx.to(dev), [n_obs] * B, times, targets, 0.05)
    assert torch.equal(sel, got), _rel(got, sel)


def test_c_program_drives_the_rollout_from_the_header(tmp_path):
    """tests/c_host/rollout_host.c -- a C99 program that includes only include/sf_b200.h and the CUDA runtime API -- reads weights,
    observations, noise and stamps from a file, runs query -> create -> schedule -> rollout -> read-back, and writes the selected
    latents; they equal the Python module's output bit for bit."""
    from streamingflow_b200.layers.temporal_ode_bayes import NNFOwithBayesianJumps

    src = os.path.join(ROOT, "tests", "c_host", "rollout_host.c")
    exe = str(tmp_path / "rollout_host")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"), src, "-o", exe,
           "-L", os.path.join(ROOT, "streamingflow_b200"), "-l:libsf_b200.so", "-L", os.path.join(cuda, "lib64"), "-lcudart",
           "-Wl,-rpath," + os.path.join(ROOT, "streamingflow_b200"), "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    dev = torch.device("cuda", 0)
    Cc, h, w, B, seed = 64, 20, 13, 2, 17
    times, targets, hx = _case(Cc, h, w, B, seed)
    m = NNFOwithBayesianJumps(Cc, Cc, make_cfg(Cc)).eval()
    sd = so.recipe_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed, 1.0)
    m.load_state_dict(sd, strict=True)
    plan = cpack.plan_rollout(times, targets, 0.05, True, "euler")
    tape = torch.randn(plan.info["n_eps"], Cc, h, w, generator=torch.Generator().manual_seed(seed + 1))
    hot = {k: v for k, v in sd.items() if k.startswith(("gru_c.", "gru_obs.", "p_model.")) and v.is_floating_point()}
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:      # int32 header, doubles, then named tensors: see rollout_host.c
        np.array([Cc, h, w, B, len(times[0]), len(targets[0]), plan.info["n_eps"], len(hot)], dtype=np.int32).tofile(f)
        np.array(times, dtype=np.float64).tofile(f)
        np.array(targets, dtype=np.float64).tofile(f)
        for k, v in hot.items():
            name = k.encode()
            np.array([len(name), v.numel()], dtype=np.int32).tofile(f)
            f.write(name)
            v.float().contiguous().numpy().tofile(f)
        hx.numpy().tofile(f)
        tape.numpy().tofile(f)
    r = subprocess.run([exe, inp, outp], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = torch.from_numpy(np.fromfile(outp, dtype=np.float32)).view(B, len(targets[0]), Cc, h, w)
    m = m.to(dev)
    m.cuda_graph = False
    m._draw_noise = lambda n, hh, ww, device: tape[:max(n, 1)].to(device).contiguous()
    with torch.no_grad():
        _, sel = m._integrate_impl(hx.to(dev), [len(times[0])] * B, times, targets, 0.05)
    assert torch.equal(sel.cpu(), got), _rel(got, sel.cpu())
    print(r.stdout.strip(), file=sys.stderr)
