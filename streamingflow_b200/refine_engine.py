"""Post-ODE refinement on the conv-stage kernels ("next" row 2 of SURVEY.md 8f): the two SpatialGRUs, the ConvNeXt Block
and the DeepLabHead that FuturePredictionODE.forward applies to the decoded frames (future_prediction_ode.py:56-62).

  SpatialGRU (layers/temporal.py:26-57)  T sequential ConvGRU cells on [B, C, H, W]: per step the `gates` stage (one u|r pair,
        flag SINGLE) and the `propose` stage of the ODE cell's kernels (the new state overwrites the fp32 / bf16 state in
        place), then the 1x1 `conv_decoder` as a bias_act stage whose A operand is the state and whose output image is
        frame (b, t).  The per-sample image indirection of the event table (sample id vs x image) does the (b, t) addressing.
  Block (convolutions.py:310-346)        depthwise 7x7 + LayerNorm in one SIMT kernel (sf_dwconv7_ln), Linear 64->256 + GELU and
        Linear 256->64 (+ layer scale folded into the weights) + residual as 1x1 conv stages.
  DeepLabHead (convolutions.py:198-250)  ASPP: 1x1 and three dilated 3x3 branches (dilated taps = 1x1 chunks with shifted input
        windows), the image-pooling branch folded into a per-image bias of the 640->128 projection (sf_aspp_pool_bias), then
        3x3 conv + BN + ReLU and the final 1x1.  BatchNorms are folded (eval mode).
All activations are NHWC bf16 planes of ONE plan at the BEV resolution with n = B*T images; the input planes are the fused
decoder's output buffer and the result leaves as NCHW fp32.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List

import numpy as np
import torch

from . import _lib as L
from .codec_engine import LevelPlan, _act_flags
from .engine import StageDef

(R_X, R_S, R_U, R_G, R_O0, R_DW, R_P1, R_BK, R_O1, R_A0, R_A1, R_A2, R_A3, R_PR, R_H, R_OUT) = range(16)
def frame_bufs(Cc: int):
    """Channel counts of the per-frame buffers; Cc = FuturePredictionODE's in_channels (64, or 128 for BASELINE config 5)."""
    return {R_X: Cc, R_O0: Cc, R_DW: Cc, R_P1: 4 * Cc, R_BK: Cc, R_O1: Cc, R_A0: 128, R_A1: 128, R_A2: 128, R_A3: 128, R_PR: 128, R_H: 128, R_OUT: Cc}


def state_bufs(Cc: int):
    return {R_S: Cc, R_U: Cc, R_G: Cc}


FRAME_BUFS, STATE_BUFS = frame_bufs(64), state_bufs(64)
ASPP_RATES = (12, 24, 36)


def _fold(w, sd, bn):
    scale = sd[bn + ".weight"].float() / torch.sqrt(sd[bn + ".running_var"].float() + 1e-5)
    return w.float() * scale[:, None, None, None], sd[bn + ".bias"].float() - sd[bn + ".running_mean"].float() * scale


def refine_graph(sd: Dict[str, torch.Tensor], fuse_pw: bool = False):
    """Stage definitions and kernel parameters of the refinement, from FuturePredictionODE's state_dict (keys spatial_grus.*,
    res_blocks.*).  Returns a dict consumed by RefineEngine and by the CPU interpreter in tests/.
    fuse_pw (64 channels, bf16): the ConvNeXt block's pwconv1 -> GELU -> pwconv2 -> layer scale -> + residual as ONE stage whose
    4C-channel intermediate stays in tensor memory (stage flag FLAG_PW_B2B) instead of three launches around a [n, H, W, 4C] buffer."""
    g = {}
    Cc = g["C"] = sd["spatial_grus.0.conv_update.weight"].shape[0]          # SpatialGRU(in_channels, in_channels): weights [Cc, 2 Cc, 3, 3]
    for i, (src, dst) in enumerate(((R_X, R_O0), (R_BK, R_O1))):
        p = f"spatial_grus.{i}."
        wu, wr, wt = (sd[p + k + ".weight"].float() for k in ("conv_update", "conv_reset", "conv_state_tilde"))
        gates = StageDef(f"gru{i}.gates", L.EPI_GATES, torch.cat([sd[p + "conv_update.bias"].float(), sd[p + "conv_reset.bias"].float()]),
                         [R_U, R_G], flags=L.FLAG_SINGLE)
        gates.add(R_S, torch.cat([wu[:, Cc:], wr[:, Cc:]], 0), 0, 1).add(L.SRC_X, torch.cat([wu[:, :Cc], wr[:, :Cc]], 0), 0, 0)
        prop = StageDef(f"gru{i}.propose", L.EPI_PROPOSE, sd[p + "conv_state_tilde.bias"].float(), [R_U, R_S], flags=L.FLAG_SINGLE | L.FLAG_KEEP_A32)
        prop.add(L.SRC_X, wt[:, :Cc], 0, 1).add(R_G, wt[:, Cc:], 0, 0)
        dec = StageDef(f"gru{i}.dec", L.EPI_BIAS_LRELU, torch.zeros(Cc, device=wu.device), [dst], flags=_act_flags(L.ACT_NONE))
        dec.add(L.SRC_X, sd[p + "conv_decoder.weight"].float(), 0, 1)
        g[f"gru{i}"] = dict(src=src, dst=dst, gates=gates, propose=prop, dec=dec)
    b = "res_blocks.0.0."
    w1, b1 = sd[b + "pwconv1.weight"].float(), sd[b + "pwconv1.bias"].float()          # Linear [4 Cc, Cc]
    w2, b2 = sd[b + "pwconv2.weight"].float(), sd[b + "pwconv2.bias"].float()          # Linear [Cc, 4 Cc]
    gamma = sd[b + "gamma"].float() if (b + "gamma") in sd else torch.ones_like(b2)
    blk = dict(dw_w=sd[b + "dwconv.weight"].float().contiguous(), dw_b=sd[b + "dwconv.bias"].float().contiguous(),
               ln_w=sd[b + "norm.weight"].float().contiguous(), ln_b=sd[b + "norm.bias"].float().contiguous(), stages=[])
    if fuse_pw:
        assert Cc == 64, "the fused pointwise pair is built for 64 channels"
        st = StageDef("block.pw", L.EPI_RES_ID, torch.cat([b1, gamma * b2]), [R_O0, R_BK], flags=_act_flags(L.ACT_NONE) | L.FLAG_PW_B2B)
        st.add(R_DW, w1[:, :, None, None], 0, 1)
        # pwconv2 [C n][4C k] as four K-chunks of [C n][64 k] rows: the B operand tiles of the back-to-back GEMM
        st.b2b_w = (gamma[:, None] * w2).view(Cc, 4 * Cc // 64, 64).permute(1, 0, 2).reshape(4 * Cc, 64).contiguous()
        blk["stages"].append(st)
    else:
        for h in range(4 * Cc // 128):
            r = slice(128 * h, 128 * h + 128)
            blk["stages"].append(StageDef(f"block.pw1{'abcd'[h]}", L.EPI_BIAS_LRELU, b1[r], [R_P1], [128 * h], flags=_act_flags(L.ACT_GELU))
                                 .add(R_DW, w1[r][:, :, None, None], 0, 1))
        blk["stages"].append(StageDef("block.pw2", L.EPI_RES_ID, gamma * b2, [R_O0, R_BK], flags=_act_flags(L.ACT_NONE))
                             .add(R_P1, (gamma[:, None] * w2)[:, :, None, None], 0, 1))
    g["block"] = blk
    d = "res_blocks.1."
    a = d + "0."
    st = []
    w, bb = _fold(sd[a + "convs.0.0.weight"], sd, a + "convs.0.1")
    st.append(StageDef("aspp.b0", L.EPI_BIAS_LRELU, bb, [R_A0], flags=_act_flags(L.ACT_RELU)).add(R_O1, w, 0, 1))
    for k, rate in enumerate(ASPP_RATES, start=1):
        w, bb = _fold(sd[f"{a}convs.{k}.0.weight"], sd, f"{a}convs.{k}.1")
        st.append(StageDef(f"aspp.b{k}", L.EPI_BIAS_LRELU, bb, [R_A0 + k], flags=_act_flags(L.ACT_RELU)).add_dilated(R_O1, w, 0, 1, rate))
    wpool, bpool = _fold(sd[a + "convs.4.1.weight"], sd, a + "convs.4.2")
    wproj, bproj = _fold(sd[a + "project.0.weight"], sd, a + "project.1")              # [128, 640, 1, 1]
    g["pool"] = dict(pool_w=wpool[:, :, 0, 0].contiguous(), pool_b=bpool.contiguous(), proj_w=wproj[:, 512:640, 0, 0].contiguous(),
                     proj_b=bproj.contiguous())          # pool_w [128, Cc]
    proj = StageDef("aspp.project", L.EPI_BIAS_LRELU, torch.zeros(128, device=wproj.device), [R_PR], flags=_act_flags(L.ACT_RELU) | L.FLAG_IMG_BIAS)
    for k in range(4):
        proj.add(R_A0 + k, wproj[:, 128 * k:128 * k + 128], 0, k == 0)
    st.append(proj)
    w, bb = _fold(sd[d + "1.weight"], sd, d + "2")
    st.append(StageDef("head.conv3", L.EPI_BIAS_LRELU, bb, [R_H], flags=_act_flags(L.ACT_RELU)).add(R_PR, w, 0, 1))
    st.append(StageDef("head.out", L.EPI_BIAS_LRELU, sd[d + "4.bias"].float(), [R_OUT], flags=_act_flags(L.ACT_NONE, out32=True))
              .add(R_H, sd[d + "4.weight"].float(), 0, 1))
    g["deeplab"] = st
    return g


class RefineEngine:
    def __init__(self, sd: Dict[str, torch.Tensor], H: int, W: int, B: int, T: int, precision: str, device):
        self.lib = L.load()
        self.H, self.W, self.B, self.T, self.device = H, W, B, T, device
        self.x3 = precision == "bf16x3"
        n = B * T
        sd = {k: v.detach().to(device) for k, v in sd.items() if k.startswith(("spatial_grus", "res_blocks"))}
        Cc = self.C = sd["spatial_grus.0.conv_update.weight"].shape[0]
        if Cc not in (64, 128) or sd["spatial_grus.0.conv_update.weight"].shape[1] != 2 * Cc or "res_blocks.0.1.dwconv.weight" in sd \
                or sd["res_blocks.1.0.convs.0.0.weight"].shape[0] != 128:
            raise L.SfError("fused refinement is built for 64 or 128 channels, one ConvNeXt block and DeepLabHead(C, C, 128)")
        # 64 channels, bf16: the block's pointwise pair runs as one back-to-back GEMM stage (SF_PW_B2B=0: three launches)
        self.fuse_pw = Cc == 64 and not self.x3 and os.environ.get("SF_PW_B2B", "1") != "0"
        self.g = refine_graph(sd, fuse_pw=self.fuse_pw)
        P = self.plan = LevelPlan(self.lib, H, W, n, self.x3, device, C_hidden=Cc)
        for b, ch in frame_bufs(Cc).items():
            if b != R_X and not (b == R_P1 and self.fuse_pw):          # the fused pair has no 4C-channel intermediate in HBM
                P.buf(b, ch)
        for b, ch in state_bufs(Cc).items():
            shape = (B, H, W, ch)
            P.buf(b, ch, planes=(torch.zeros(shape, dtype=torch.bfloat16, device=device),
                                 torch.zeros(shape, dtype=torch.bfloat16, device=device) if self.x3 else None))
        self.state32 = torch.zeros((B, H, W, Cc), dtype=torch.float32, device=device)
        for slot in (L.F32_STATE0, L.F32_A):
            L.check(self.lib.sf_plan_bind_f32(P.plan, slot, self.state32.data_ptr()), "bind state")
        self.out32 = torch.empty((n, H, W, Cc), dtype=torch.float32, device=device)
        P.bind_out32(self.out32)
        self.img_bias = torch.zeros((n, 128), dtype=torch.float32, device=device)
        L.check(self.lib.sf_plan_bind_f32(P.plan, L.F32_IMG_BIAS, self.img_bias.data_ptr()), "bind img bias")
        self.pool_scratch = torch.zeros((n, 32, Cc), dtype=torch.float32, device=device)
        self.slots = {}
        for i in range(2):
            gi = self.g[f"gru{i}"]
            self.slots[f"gru{i}"] = tuple(P.stage(gi[k]) for k in ("gates", "propose", "dec"))
        self.slots["block"] = [P.stage(s) for s in self.g["block"]["stages"]]
        self.slots["deeplab"] = [P.stage(s) for s in self.g["deeplab"]]
        P.finalize()
        # event tables: frame-parallel stages use LevelPlan.run; the GRU steps need (sample id, x image) pairs
        rows, self.ev_gru, self.ev_dec, off = [], {}, {}, 0
        for t in range(T):
            frames = np.arange(B) * T + t
            for name, sid, ximg in (("gru", np.arange(B), frames), ("dec", frames, np.arange(B))):
                blk = np.zeros((5, B), dtype=np.int32)
                blk[0], blk[1], blk[2] = sid, ximg, -1
                rows.append(blk.reshape(-1))
                (self.ev_gru if name == "gru" else self.ev_dec)[t] = off
                off += 5 * B
        self.table = torch.from_numpy(np.concatenate(rows)).to(device)
        self._first_frames = torch.arange(B, device=device) * T
        self.launches = 0

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _run_gru(self, i: int, x_planes):
        P, lib, B, T = self.plan, self.lib, self.B, self.T
        gi = self.g[f"gru{i}"]
        s_gates, s_prop, s_dec = self.slots[f"gru{i}"]
        stream = self._stream()
        with torch.cuda.device(self.device):
            for t in range(T):
                ev = L.Event(0, B, gi["src"], 0, 0, 0, 1, 0, 0, self.ev_gru[t])
                L.check(lib.sf_plan_run_stage(P.plan, s_gates, C.byref(ev), self.table.data_ptr(), stream), "gru gates")
                L.check(lib.sf_plan_run_stage(P.plan, s_prop, C.byref(ev), self.table.data_ptr(), stream), "gru propose")
                ev2 = L.Event(0, B, R_S, 0, 0, 0, 1, 0, 0, self.ev_dec[t])
                L.check(lib.sf_plan_run_stage(P.plan, s_dec, C.byref(ev2), self.table.data_ptr(), stream), "gru dec")
        self.launches += 3 * T

    def _init_state(self, x_planes, x32):
        B, T = self.B, self.T
        idx = self._first_frames                                  # hidden_state = x[:, 0] (future_prediction_ode.py:56)
        self.state32.copy_(x32.view(B * T, self.H, self.W, self.C)[idx])
        for dst, src in zip(self.plan.bufs[R_S], x_planes):
            if dst is not None:
                dst.copy_(src[idx])

    def output_planes(self):
        """The refined frames in engine layout: (hi, lo) NHWC bf16 [B*T, H, W, C] (image index b*T + t)."""
        return self.plan.bufs[R_OUT]

    def run(self, x_planes, x32: torch.Tensor) -> torch.Tensor:
        """x_planes: (hi, lo) NHWC bf16 [B*T, H, W, C] decoded frames, image index b*T + t; x32: the same frames in fp32 NHWC.
        Returns the refined frames [B, T, C, H, W] fp32 NCHW."""
        self.run_core(x_planes, x32)
        return self.unpack_output()

    def run_core(self, x_planes, x32: torch.Tensor):
        """Everything up to the head's fp32 NHWC output buffer (``out32``) and the bf16 output planes: static buffers in, static
        buffers out, no allocation the caller keeps -- the part a CUDA graph of the whole forward captures."""
        P, lib, n = self.plan, self.lib, self.B * self.T
        P.buf(R_X, self.C, planes=(x_planes[0][:n], x_planes[1][:n] if x_planes[1] is not None else None))
        stream = self._stream()
        self._init_state(x_planes, x32)
        self._run_gru(0, x_planes)
        blk = self.g["block"]
        o0, dw = P.bufs[R_O0], P.bufs[R_DW]
        L.check(lib.sf_dwconv7_ln(o0[0].data_ptr(), o0[1].data_ptr() if o0[1] is not None else None, dw[0].data_ptr(),
                                  dw[1].data_ptr() if dw[1] is not None else None, blk["dw_w"].data_ptr(), blk["dw_b"].data_ptr(),
                                  blk["ln_w"].data_ptr(), blk["ln_b"].data_ptr(), n, self.C, self.H, self.W, stream), "dwconv7_ln")
        self.launches += 1 + P.run(self.slots["block"], n)
        self._init_state(x_planes, x32)
        self._run_gru(1, x_planes)
        pl, o1 = self.g["pool"], P.bufs[R_O1]
        L.check(lib.sf_aspp_pool_bias(o1[0].data_ptr(), o1[1].data_ptr() if o1[1] is not None else None, pl["pool_w"].data_ptr(),
                                      pl["pool_b"].data_ptr(), pl["proj_w"].data_ptr(), pl["proj_b"].data_ptr(), self.pool_scratch.data_ptr(),
                                      self.img_bias.data_ptr(), n, self.C, self.H, self.W, stream), "aspp_pool_bias")
        self.launches += 2 + P.run(self.slots["deeplab"], n)

    def unpack_output(self) -> torch.Tensor:
        """out32 (NHWC fp32) -> a fresh [B, T, C, H, W] fp32 NCHW tensor."""
        n = self.B * self.T
        out = torch.empty((n, self.C, self.H, self.W), dtype=torch.float32, device=self.device)
        L.check(self.lib.sf_unpack_nhwc_f32(self.out32.data_ptr(), out.data_ptr(), None, n, self.C, self.H, self.W, self._stream()), "unpack")
        self.launches += 1
        return out.view(self.B, self.T, self.C, self.H, self.W)
