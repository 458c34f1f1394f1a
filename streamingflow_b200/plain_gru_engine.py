"""The reference's plain ConvGRU cells on the conv-stage kernels: ``SpatialGRUODECell`` and
``SpatialGRUCell`` (streamingflow/layers/temporal_ode_bayes.py:14-61, 165-208) -- defined by the reference but not wired into
its model.  One evaluation = the `gates` stage (one update / reset pair) and the `propose` stage of the ODE cell's kernels,
with the proposal's BatchNorm folded into the conv and its ReLU + the blend (or the GRU-ODE derivative u (s~ - s)) in the
epilogue.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .codec_engine import LevelPlan, _act_flags
from .engine import StageDef

B_X, B_S, B_U, B_G, B_OUT = range(5)


class PlainGruEngine:
    def __init__(self, sd, gru_bias_init: float, H: int, W: int, n: int, precision: str, device, deriv: bool):
        self.lib = L.load()
        self.H, self.W, self.n, self.device = H, W, n, device
        x3 = precision == "bf16x3"
        sd = {k: v.detach().to(device).float() for k, v in sd.items() if v.is_floating_point()}
        wu, wr, wt = sd["conv_update.weight"], sd["conv_reset.weight"], sd["conv_state_tilde.conv.weight"]
        if tuple(wu.shape[:2]) != (64, 128):
            raise L.SfError("the CUDA ConvGRU cells are built for 64 input and 64 hidden channels")
        scale = sd["conv_state_tilde.norm.weight"] / torch.sqrt(sd["conv_state_tilde.norm.running_var"] + 1e-5)
        wt = wt * scale[:, None, None, None]
        bt = sd["conv_state_tilde.norm.bias"] - sd["conv_state_tilde.norm.running_mean"] * scale
        if "conv_state_tilde.conv.bias" in sd:
            bt = bt + sd["conv_state_tilde.conv.bias"] * scale
        P = self.plan = LevelPlan(self.lib, H, W, n, x3, device)
        for b in (B_X, B_S, B_U, B_G, B_OUT):
            P.buf(b, 64)
        self.s32 = torch.zeros((n, H, W, 64), dtype=torch.float32, device=device)
        self.out32 = torch.zeros((n, H, W, 64), dtype=torch.float32, device=device)
        L.check(self.lib.sf_plan_bind_f32(P.plan, L.F32_STATE0, self.s32.data_ptr()), "bind state")
        L.check(self.lib.sf_plan_bind_f32(P.plan, L.F32_A, self.out32.data_ptr()), "bind out")
        gates = StageDef("gates", L.EPI_GATES, torch.cat([sd["conv_update.bias"], sd["conv_reset.bias"]]) + float(gru_bias_init), [B_U, B_G],
                         flags=L.FLAG_SINGLE)
        gates.add(B_S, torch.cat([wu[:, 64:], wr[:, 64:]], 0), 0, 1).add(L.SRC_X, torch.cat([wu[:, :64], wr[:, :64]], 0), 0, 0)
        prop = StageDef("propose", L.EPI_PROPOSE, bt, [B_U, B_OUT],
                        flags=L.FLAG_SINGLE | L.FLAG_KEEP_A32 | _act_flags(L.ACT_RELU) | (L.FLAG_DERIV if deriv else 0))
        prop.add(L.SRC_X, wt[:, :64], 0, 1).add(B_G, wt[:, 64:], 0, 0)
        self.slots = [P.stage(gates), P.stage(prop)]
        P.finalize()
        self.table = torch.zeros(5 * n, dtype=torch.int32, device=device)
        self.table[:n] = torch.arange(n, dtype=torch.int32, device=device)
        self.table[n:2 * n] = torch.arange(n, dtype=torch.int32, device=device)
        self.table[2 * n:3 * n] = -1

    def run(self, x: torch.Tensor, state: torch.Tensor) -> torch.Tensor:
        n = x.shape[0]
        assert n == self.n and tuple(x.shape[1:]) == (64, self.H, self.W) and x.shape == state.shape
        lib, P = self.lib, self.plan
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        ptr = lambda t: t.data_ptr() if t is not None else None
        for buf, t in ((B_X, x), (B_S, state)):
            src = t.contiguous().float()
            hi, lo = P.bufs[buf]
            L.check(lib.sf_pack_nchw_f32(src.data_ptr(), hi.data_ptr(), ptr(lo), n, 64, self.H, self.W, stream), "pack")
        self.s32.copy_(state.float().permute(0, 2, 3, 1))
        ev = L.Event(0, n, B_X, 0, 0, 0, 1, 0, 0, 0)
        with torch.cuda.device(self.device):
            for s in self.slots:
                L.check(lib.sf_plan_run_stage(P.plan, s, C.byref(ev), self.table.data_ptr(), stream), "run_stage")
        out = torch.empty((n, 64, self.H, self.W), dtype=torch.float32, device=self.device)
        L.check(lib.sf_unpack_nhwc_f32(self.out32.data_ptr(), out.data_ptr(), None, n, 64, self.H, self.W, stream), "unpack")
        return out
