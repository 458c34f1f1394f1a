"""Host-side driver of the CUDA ODE engine: packs the reference's weights into the kernels' layouts,
owns the NHWC workspace, and turns per-sample step schedules into batched event launches through the
C ABI (include/sf_b200.h).  Python / PyTorch here is plumbing only: device memory, streams, RNG.

Weight sources (reference parameter names) -> conv stages:
  gates    conv_update_1/2, conv_reset_1/2            temporal_ode_bayes.py:135-140,150-155
  propose  conv_state_tilde_1/2 + GRU blend           :143-146,158-161
  decode   conv_decoder_2                             :121
  trunk7   trusting_gate.0.layers.0 (7x7) + LN + GELU convolutions.py:356-358
  trunk1   layers.3 (1x1) + LN + GELU                 :359-361
  mix      layers.6 (3x3)+LN+GELU, projection, 1x1->2, softmax, mix, Euler/jump   :362-380, tob:124-131,446
  q1..q5   p_model ConvNet with BatchNorm folded      res_models.py:168-180
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

# activation buffer ids
(BUF_S0, BUF_S1, BUF_X, BUF_OBS, BUF_ZERO, BUF_U1, BUF_U2, BUF_G1, BUF_G2, BUF_A, BUF_B, BUF_HH, BUF_T1, BUF_T2, BUF_Q1,
 BUF_Z1, BUF_Y1, BUF_Q3, BUF_Z2, BUF_Y2) = range(20)
_BUF_CHANNELS = {BUF_S0: 1, BUF_S1: 1, BUF_X: 1, BUF_U1: 1, BUF_U2: 1, BUF_G1: 1, BUF_G2: 1, BUF_A: 1, BUF_B: 1, BUF_HH: 1,
                 BUF_T1: 1, BUF_T2: 1, BUF_Q1: 1, BUF_Z1: 2, BUF_Y1: 2, BUF_Q3: 2, BUF_Z2: 2, BUF_Y2: 2}
KIND_STEP, KIND_JUMP = 0, 1
CELL_STAGE_NAMES = {64: ("gates", "propose", "decode", "trunk", "mix"),        # SF_B2B=0: "trunk" splits into "trunk7", "trunk1"
                    128: ("gates_1", "gates_2", "propose_1", "propose_2", "decode", "trunk7", "trunk1", "mix")}


def _bn_fold(sd, p):
    """eval-mode BatchNorm2d folded into the preceding bias-free conv: w' = w * g/sqrt(var+1e-5), b' = beta - mean*scale."""
    w = sd[p + ".conv.weight"].float()
    scale = sd[p + ".norm.weight"].float() / torch.sqrt(sd[p + ".norm.running_var"].float() + 1e-5)
    bias = sd[p + ".norm.bias"].float() - sd[p + ".norm.running_mean"].float() * scale
    return w * scale[:, None, None, None], bias


class StageDef:
    """Logical description of a stage before packing: chunks of (source buffer, first channel, weights [n,64,R,R], column, init)
    plus the epilogue's io buffers (with the first channel each launch touches) and flags."""

    def __init__(self, name, epilogue, vec, io, io_off=None, flags=0):
        self.name, self.epilogue, self.vec, self.io, self.flags = name, epilogue, vec, list(io), flags
        self.io_off = list(io_off) if io_off is not None else [0] * len(self.io)
        self.chunks: List[Tuple[int, int, torch.Tensor, int, int, int, int]] = []
        self.fold_se: Optional[int] = None      # SE layer whose per-sample scales are folded into this stage's weights
        self.b2b_w: Optional[torch.Tensor] = None   # [C, 64] weights of a fused 1x1 follow-up conv (FLAG_B2B)

    def add(self, buf, w, col, init, c0=0, ox=0, oy=0):
        """w: [n, cin, R, R] with cin a multiple of 64: one chunk per 64 input channels of buffer ``buf`` starting at c0.
        (ox, oy) shifts the chunk's input window (used for dilated taps)."""
        assert w.shape[1] % 64 == 0 and w.shape[2] == w.shape[3]
        for i in range(w.shape[1] // 64):
            self.chunks.append((buf, c0 + 64 * i, w[:, 64 * i:64 * i + 64], col, int(init and i == 0), ox, oy))
        return self

    def add_dilated(self, buf, w, col, init, dilation):
        """3x3 convolution with dilation d (padding d): nine 1x1 chunks whose windows are shifted by ((kx-1) d, (ky-1) d)."""
        assert w.shape[2] == 3 and w.shape[3] == 3
        first = True
        for ky in range(3):
            for kx in range(3):
                self.add(buf, w[:, :, ky:ky + 1, kx:kx + 1], col, init and first, ox=(kx - 1) * dilation, oy=(ky - 1) * dilation)
                first = False
        return self


def cell_stage_defs(sd: Dict[str, torch.Tensor], p: str, pair_rows: bool = True, b2b: bool = True, pair3: bool = False) -> List[StageDef]:
    """The conv stages of one dual-GRU cell (derivative or jump) from the reference's parameters; C = 64 or 128.
    b2b (C = 64): the Bottleblock's 7x7 conv + LN + GELU + 1x1 conv + LN + GELU run as ONE stage ("trunk")."""
    g = lambda k: sd[f"{p}.{k}"].float()
    C_ = g("conv_update_1.weight").shape[0]
    assert C_ in (64, 128)
    wu1, wr1, wt1 = g("conv_update_1.weight"), g("conv_reset_1.weight"), g("conv_state_tilde_1.weight")
    wu2, wr2, wt2 = g("conv_update_2.weight"), g("conv_reset_2.weight"), g("conv_state_tilde_2.weight")
    bu1, br1, bu2, br2 = g("conv_update_1.bias"), g("conv_reset_1.bias"), g("conv_update_2.bias"), g("conv_reset_2.bias")
    fold = lambda w: w[:, :C_] + w[:, C_:]        # gru_cell_2 sees cat[state, state] (tob:118): one C-channel operand
    SX, SS = L.SRC_X, L.SRC_STATE_IN
    out = []
    if C_ == 64:        # all four gates in one 256-column launch
        st = StageDef("gates", L.EPI_GATES, torch.cat([bu1, br1, bu2, br2]), [BUF_U1, BUF_G1, BUF_U2, BUF_G2])
        st.add(SS, torch.cat([wu1[:, C_:], wr1[:, C_:], fold(wu2), fold(wr2)], 0), 0, 1).add(SX, torch.cat([wu1[:, :C_], wr1[:, :C_]], 0), 0, 0)
        out.append(st)
        st = StageDef("propose", L.EPI_PROPOSE, torch.cat([g("conv_state_tilde_1.bias"), g("conv_state_tilde_2.bias")]),
                      [BUF_U1, BUF_U2, BUF_A, BUF_HH], flags=1)
        st.add(SX, wt1[:, :C_], 0, 1).add(BUF_G1, wt1[:, C_:], 0, 0).add(SS, wt2[:, :C_], 64, 1).add(BUF_G2, wt2[:, C_:], 64, 0)
        out.append(st)
    else:               # 128 channels: one gate pair / one proposal per launch
        st = StageDef("gates_1", L.EPI_GATES, torch.cat([bu1, br1]), [BUF_U1, BUF_G1])
        st.add(SS, torch.cat([wu1[:, C_:], wr1[:, C_:]], 0), 0, 1).add(SX, torch.cat([wu1[:, :C_], wr1[:, :C_]], 0), 0, 0)
        out.append(st)
        out.append(StageDef("gates_2", L.EPI_GATES, torch.cat([bu2, br2]), [BUF_U2, BUF_G2]).add(SS, torch.cat([fold(wu2), fold(wr2)], 0), 0, 1))
        out.append(StageDef("propose_1", L.EPI_PROPOSE, g("conv_state_tilde_1.bias"), [BUF_U1, BUF_A], flags=1)
                   .add(SX, wt1[:, :C_], 0, 1).add(BUF_G1, wt1[:, C_:], 0, 0))
        out.append(StageDef("propose_2", L.EPI_PROPOSE, g("conv_state_tilde_2.bias"), [BUF_U2, BUF_HH])
                   .add(SS, wt2[:, :C_], 0, 1).add(BUF_G2, wt2[:, C_:], 0, 0))
    out.append(StageDef("decode", L.EPI_DECODE, g("conv_decoder_2.bias"), [BUF_B], flags=L.FLAG_PAIR_ROWS if (C_ == 64 and pair3) else 0)
               .add(BUF_HH, g("conv_decoder_2.weight"), 0, 1))
    t = "trusting_gate.0."
    w7 = g(t + "layers.0.weight")
    # 64 channels: vertically adjacent taps of the 7x7 conv are paired into N = 128 MMAs (the MMA issue cost is ~41 + N/2 cycles)
    if C_ == 64 and b2b:
        st = StageDef("trunk", L.EPI_LNGELU, torch.cat([g(t + "layers.1.weight"), g(t + "layers.1.bias"), g(t + "layers.4.weight"),
                                                        g(t + "layers.4.bias")]), [BUF_T2],
                      flags=L.FLAG_B2B | (L.FLAG_PAIR_ROWS if pair_rows else 0))
        st.add(BUF_A, w7[:, :C_], 0, 1).add(BUF_B, w7[:, C_:], 0, 0)
        st.b2b_w = g(t + "layers.3.weight")[:, :, 0, 0]
        out.append(st)
    else:
        out.append(StageDef("trunk7", L.EPI_LNGELU, torch.cat([g(t + "layers.1.weight"), g(t + "layers.1.bias")]), [BUF_T1],
                            flags=L.FLAG_PAIR_ROWS if (C_ == 64 and pair_rows) else 0)
                   .add(BUF_A, w7[:, :C_], 0, 1).add(BUF_B, w7[:, C_:], 0, 0))
        out.append(StageDef("trunk1", L.EPI_LNGELU, torch.cat([g(t + "layers.4.weight"), g(t + "layers.4.bias")]), [BUF_T2])
                   .add(BUF_T1, g(t + "layers.3.weight"), 0, 1))
    wp = g(t + "projection.0.weight")
    wg = g("trusting_gate.1.weight")[:, :, 0, 0]
    out.append(StageDef("mix", L.EPI_MIX, torch.cat([g(t + "layers.7.weight"), g(t + "layers.7.bias"), wg[0], wg[1]]), [])
               .add(BUF_T2, g(t + "layers.6.weight"), 0, 1).add(BUF_A, wp[:, :C_], C_, 1).add(BUF_B, wp[:, C_:], C_, 0))
    return out


def prior_stage_defs(sd: Dict[str, torch.Tensor], p: str, fold_se: bool = False, pair3: bool = False) -> List[object]:
    """p_model = ConvNet(C, 2C) with BatchNorm folded (res_models.py:168-180), as a list of StageDefs and the two SE markers
    'se0' / 'se1'.  Outputs wider than 128 channels are produced 128 channels per launch.

    fold_se: the SE layers (res_models.py:150-165) are not applied to the activation tensors; their per-sample channel scales
    are folded into the weights of the convs that consume them (q3, q5) and into q4's residual, which then read z itself."""
    Y1, Y2 = (BUF_Z1, BUF_Z2) if fold_se else (BUF_Y1, BUF_Y2)
    m = p + ".model."
    w1, b1 = _bn_fold(sd, m + "0.layers.conv_1")
    w2, b2 = _bn_fold(sd, m + "0.layers.conv_2")
    C_ = w1.shape[0]
    halves = (2 * C_) // 128
    SO = L.SRC_STATE_OUT
    items: List[object] = [StageDef("q1", L.EPI_BIAS_LRELU, b1, [BUF_Q1], flags=L.FLAG_PAIR_ROWS if (C_ == 64 and pair3) else 0).add(SO, w1, 0, 1)]
    wpj, bpj = sd[m + "0.projection.weight"].float(), sd[m + "0.projection.bias"].float()
    for h in range(halves):
        r = slice(128 * h, 128 * h + 128)
        items.append(StageDef(f"q2{'ab'[h] if halves > 1 else ''}", L.EPI_RES_PROJ, torch.cat([b2[r], bpj[r]]), [BUF_Z1], [128 * h])
                     .add(BUF_Q1, w2[r], 0, 1).add(SO, wpj[r], 128, 1))
    items.append("se0")
    w3, b3 = _bn_fold(sd, m + "2.layers.conv_1")
    w4, b4 = _bn_fold(sd, m + "2.layers.conv_2")
    for h in range(halves):
        r = slice(128 * h, 128 * h + 128)
        items.append(StageDef(f"q3{'ab'[h] if halves > 1 else ''}", L.EPI_BIAS_LRELU, b3[r], [BUF_Q3], [128 * h]).add(Y1, w3[r], 0, 1))
        items[-1].fold_se = 0 if fold_se else None
    for h in range(halves):
        r = slice(128 * h, 128 * h + 128)
        items.append(StageDef(f"q4{'ab'[h] if halves > 1 else ''}", L.EPI_RES_ID, b4[r], [Y1, BUF_Z2], [128 * h, 128 * h],
                              flags=L.FLAG_RES_SE_SCALE if fold_se else 0).add(BUF_Q3, w4[r], 0, 1))
    items.append("se1")
    items.append(StageDef("q5", L.EPI_SAMPLE, sd[m + "4.conv.bias"].float(), [BUF_X]).add(Y2, sd[m + "4.conv.weight"].float(), 0, 1))
    items[-1].fold_se = 1 if fold_se else None
    return items


_PAIRABLE = None


def pair_rows_if_eligible(sdef: StageDef, C_hidden: int = 64) -> StageDef:
    """Sets FLAG_PAIR_ROWS on a stage of a 64-channel plan whose accumulator is ONE 64-column block produced by undilated 3x3 (or 7x7)
    chunks only -- the N = 64 stages that sit at the 44 % issue ceiling of an MMA with both operands in shared memory: their
    vertically adjacent taps then issue as N = 128 MMAs and the epilogue folds the second column block back one row (sf_conv.cuh
    fold_paired_rows).  Measured on B200 (profiles/r02_ab_pair3_resident.txt): a LOSS for these light-epilogue stages -- decode 37.8 ->
    45.5 us, q1 37.8 -> 41.9 us, forward 16.26 -> 16.54 ms: the fold roughly doubles their epilogue and tiles shrink to 15 rows, which
    costs more than the 19 % fewer MMA issue cycles buy (the 7x7 trunk, whose epilogue is heavy anyway, keeps its pairing).  So this
    is OFF unless SF_PAIR_3X3=1."""
    global _PAIRABLE
    if _PAIRABLE is None:
        _PAIRABLE = (L.EPI_LNGELU, L.EPI_DECODE, L.EPI_BIAS_LRELU, L.EPI_RES_ID)
    if os.environ.get("SF_PAIR_3X3", "0") != "1" or C_hidden != 64 or sdef.epilogue not in _PAIRABLE:
        return sdef
    if sdef.flags & (L.FLAG_PAIR_ROWS | L.FLAG_RES_SE_SCALE | L.FLAG_PW_B2B):
        return sdef
    if not sdef.chunks or any(w.shape[0] != 64 or w.shape[2] < 3 or col != 0 or ox or oy for _, _, w, col, _, ox, oy in sdef.chunks):
        return sdef
    sdef.flags |= L.FLAG_PAIR_ROWS
    return sdef


def pack_stage_master(sdef: StageDef, x3: bool):
    """fp32 master of pack_stage's matrix for a stage whose weights get an SE layer folded in: the same rows in the same order,
    every row holding the UNSPLIT fp32 weights, plus per-row metadata (bits 0-15: the chunk's first channel = the offset into
    the SE scale vector; bit 16: residual row of the split mode).  se_fold_kernel turns it into per-sample bf16 weights."""
    blocks, meta = [], []
    for buf, c0, w, col, init, ox, oy in sdef.chunks:
        n, _, R, _ = w.shape
        taps = w.permute(3, 2, 0, 1).contiguous().float()           # [dx, dy, n, 64]
        if not x3:
            blocks.append(taps.reshape(-1, 64))
            meta.append(torch.full((R * R * n,), c0, dtype=torch.int32))
        else:
            blocks.append(torch.stack([taps, taps], dim=2).reshape(-1, 64))      # [dx, dy, rep (hi, lo), n, 64]
            m = torch.full((R, R, 2, n), c0, dtype=torch.int32)
            m[:, :, 1] |= 1 << 16
            meta.append(m.reshape(-1))
            blocks.append(taps.reshape(-1, 64))                                   # lo-plane chunk: hi weights
            meta.append(torch.full((R * R * n,), c0, dtype=torch.int32))
    return torch.cat(blocks, 0).contiguous(), torch.cat(meta, 0).contiguous()


def _pair_order(R: int) -> List[int]:
    """dy order of a dx column with row-paired taps: [1, 0, 3, 2, ..., R-1 alone when R is odd]."""
    order = []
    for g in range(R // 2):
        order += [2 * g + 1, 2 * g]
    if R % 2:
        order.append(R - 1)
    return order


def pack_stage(sdef: StageDef, x3: bool):
    """Packs a stage's weights into the [rows, 64] bf16 matrix the TMA weight ring streams, in consumption order:
    chunk -> dx -> dy -> rep -> n rows (see sf_chunk.wrow in include/sf_b200.h).  Returns (chunks, w_packed).
    Row-paired stages (FLAG_PAIR_ROWS): chunk -> dx -> pair -> rep -> [tap_hi n rows | tap_lo n rows]."""
    if sdef.flags & (L.FLAG_B2B | L.FLAG_PW_B2B):
        plain = StageDef(sdef.name, sdef.epilogue, sdef.vec, sdef.io, sdef.io_off, sdef.flags & ~(L.FLAG_B2B | L.FLAG_PW_B2B))
        plain.chunks = sdef.chunks
        chunks, wp = pack_stage(plain, x3)
        w = sdef.b2b_w.float()                                       # [n, k] = the K-major B operand of the follow-up GEMM (64 k per row)
        hi = w.to(torch.bfloat16)
        tail = [hi] + ([(w - hi.float()).to(torch.bfloat16)] if x3 else [])
        return chunks, torch.cat([wp] + tail, 0).contiguous()
    if sdef.flags & L.FLAG_PAIR_ROWS:
        return _pack_stage_paired(sdef, x3)
    chunks, blocks, row = [], [], 0
    for buf, c0, w, col, init, ox, oy in sdef.chunks:
        n, _, R, _ = w.shape
        taps = w.permute(3, 2, 0, 1).contiguous()                  # [dx, dy, n, 64]
        hi = taps.to(torch.bfloat16)
        if not x3:
            chunks.append(dict(buf=buf, plane=0, c0=c0, R=R, n=n, nrep=1, col=col, wrow=row, init=init, ox=ox, oy=oy))
            blocks.append(hi.reshape(-1, 64))
            row += R * R * n
        else:
            lo = (taps - hi.float()).to(torch.bfloat16)
            both = torch.stack([hi, lo], dim=2)                     # [dx, dy, rep, n, 64]
            chunks.append(dict(buf=buf, plane=0, c0=c0, R=R, n=n, nrep=2, col=col, wrow=row, init=init, ox=ox, oy=oy))
            blocks.append(both.reshape(-1, 64))
            row += R * R * 2 * n
            chunks.append(dict(buf=buf, plane=1, c0=c0, R=R, n=n, nrep=1, col=col, wrow=row, init=0, ox=ox, oy=oy))
            blocks.append(hi.reshape(-1, 64))
            row += R * R * n
    return chunks, torch.cat(blocks, 0).contiguous()


def _pack_stage_paired(sdef: StageDef, x3: bool):
    chunks, blocks, row = [], [], 0
    for buf, c0, w, col, init, ox, oy in sdef.chunks:
        n, _, R, _ = w.shape
        order = _pair_order(R)
        groups = [order[i:i + 2] for i in range(0, R - R % 2, 2)] + ([[order[-1]]] if R % 2 else [])
        taps = w.permute(3, 2, 0, 1).contiguous()                  # [dx, dy, n, 64]
        hi = taps.to(torch.bfloat16)
        lo = (taps - hi.float()).to(torch.bfloat16)
        reps = [hi, lo] if x3 else [hi]
        rows = []
        for dx in range(R):
            for grp in groups:
                for plane in reps:
                    for dy in grp:
                        rows.append(plane[dx, dy])
        chunks.append(dict(buf=buf, plane=0, c0=c0, R=R, n=n, nrep=len(reps), col=col, wrow=row, init=init, ox=ox, oy=oy))
        blocks.append(torch.cat(rows, 0))
        row += R * R * len(reps) * n
        if x3:                                                      # lo plane of the activations x hi weights
            rows = [hi[dx, dy] for dx in range(R) for grp in groups for dy in grp]
            chunks.append(dict(buf=buf, plane=1, c0=c0, R=R, n=n, nrep=1, col=col, wrow=row, init=0, ox=ox, oy=oy))
            blocks.append(torch.cat(rows, 0))
            row += R * R * n
    return chunks, torch.cat(blocks, 0).contiguous()


def _emulate_paired(sdef: StageDef, x3: bool, sources: Dict[int, torch.Tensor]) -> torch.Tensor:
    """Host replay of a row-paired stage exactly as the kernel runs it: per dx column and pair, ONE product of the dy_hi window
    with the [tap_hi | tap_lo] rows; the tap_lo half lands one row above its pixel and is folded back one row down (the
    epilogue's block-1 shift), with the virtual row above the image computed like the kernel's scratch row."""
    import torch.nn.functional as F

    chunks, wp = pack_stage(sdef, x3)
    wp = wp.double()
    H, W = next(iter(sources.values())).shape[-2:]
    n = chunks[0]["n"]
    blk0 = torch.zeros(n, H + 1, W, dtype=torch.float64)           # virtual rows -1 .. H-1
    blk1 = torch.zeros(n, H + 1, W, dtype=torch.float64)
    for ck in chunks:
        x = sources[ck["buf"]][0, ck["c0"]:ck["c0"] + 64].double()
        xh = x.to(torch.bfloat16).double()
        x = (xh if ck["plane"] == 0 else (x - xh).to(torch.bfloat16).double()) if x3 else xh
        R, nrep = ck["R"], ck["nrep"]
        pad = (R - 1) // 2
        xp = F.pad(x, (pad, pad, pad + 1, pad))                      # one extra row on top: virtual row -1 reads rows -1-pad ..
        order = _pair_order(R)
        groups = [order[i:i + 2] for i in range(0, R - R % 2, 2)] + ([[order[-1]]] if R % 2 else [])
        r0 = ck["wrow"]
        for dx in range(R):
            for grp in groups:
                win = xp[:, grp[0]:grp[0] + H + 1, dx:dx + W]       # window of dy_hi for virtual rows -1 .. H-1
                for rep in range(nrep):
                    for t, dy in enumerate(grp):
                        wt = wp[r0:r0 + n]
                        r0 += n
                        prod = torch.einsum("nc,chw->nhw", wt, win)
                        (blk0 if t == 0 else blk1).add_(prod)
    acc = torch.zeros(256, H, W, dtype=torch.float64)
    acc[:n] = blk0[:, 1:] + blk1[:, :-1]                             # out[v] = block0[v] + block1[v - 1]
    return acc


def emulate_stage(sdef: StageDef, x3: bool, sources: Dict[int, torch.Tensor]) -> torch.Tensor:
    """Numerically replays the packed plan on the host (pure indexing of the packed matrix, fp64 products): the
    accumulator columns [pixels..., 256] the tensor core would produce.  Used by CPU tests to pin the packing and
    the chunk / tap / column bookkeeping against F.conv2d without a GPU.  sources[buf] is NCHW [1, Cbuf, H, W]."""
    import torch.nn.functional as F

    if sdef.flags & L.FLAG_PAIR_ROWS:
        return _emulate_paired(sdef, x3, sources)
    chunks, wp = pack_stage(sdef, x3)
    wp = wp.double()
    any_src = next(iter(sources.values()))
    H, W = any_src.shape[-2:]
    acc = torch.zeros(256, H, W, dtype=torch.float64)
    for ck in chunks:
        x = sources[ck["buf"]][0, ck["c0"]:ck["c0"] + 64].double()
        if x3:
            xh = x.to(torch.bfloat16).double()
            x = xh if ck["plane"] == 0 else (x - xh).to(torch.bfloat16).double()
        else:
            x = x.to(torch.bfloat16).double()
        R, n, nrep = ck["R"], ck["n"], ck["nrep"]
        pad = (R - 1) // 2
        ox, oy = ck.get("ox", 0), ck.get("oy", 0)
        if ox or oy:                                               # window shifted by (ox, oy), zero outside the image
            big = F.pad(x, (abs(ox), abs(ox), abs(oy), abs(oy)))
            x = big[:, abs(oy) + oy:abs(oy) + oy + H, abs(ox) + ox:abs(ox) + ox + W]
        xp = F.pad(x, (pad, pad, pad, pad))
        out = torch.zeros(n, H, W, dtype=torch.float64)
        for dx in range(R):
            for dy in range(R):
                for rep in range(nrep):
                    r0 = ck["wrow"] + ((dx * R + dy) * nrep + rep) * n
                    wt = wp[r0:r0 + n]                                             # [n, 64]
                    out += torch.einsum("nc,chw->nhw", wt, xp[:, dy:dy + H, dx:dx + W])
        if ck["init"]:
            acc[ck["col"]:ck["col"] + n] = out
        else:
            acc[ck["col"]:ck["col"] + n] += out
    return acc


class OdeEngine:
    """One plan + workspace for a fixed (max_images, H, W, precision) on one CUDA device."""

    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, H: int, W: int, max_images: int, precision: str = "bf16",
                 device: Optional[torch.device] = None, path_slots: int = 0, se_fold: bool = True):
        """se_fold: fold the two SE layers of p_model into their consumers' weights (default).  False keeps the separate
        reduce / apply kernels -- required when the channel sums are all-reduced across GPUs in between (row sharding)."""
        self.lib = L.load()
        self.se_fold = bool(se_fold)
        dev = torch.device(device if device is not None else "cuda")
        if dev.type != "cuda":
            raise L.SfError("the ODE engine runs on a CUDA (B200) device only; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.C = int(next(v for k, v in sd.items() if k.endswith("gru_c.conv_decoder_2.weight")).shape[0])
        if self.C not in (64, 128):
            raise L.SfError("the CUDA ODE engine is built for 64 or 128 hidden channels")
        self.device, self.H, self.W = dev, int(H), int(W)
        self.max_images = int(max_images)
        self.precision = precision
        self.x3 = precision == "bf16x3"
        if precision not in ("bf16", "bf16x3"):
            raise ValueError("precision must be 'bf16' or 'bf16x3'")
        L.check(self.lib.sf_device_supported(dev.index), "sf_device_supported")
        geo = L.Geometry(self.max_images, self.H, self.W, self.C, L.PREC_BF16X3 if self.x3 else L.PREC_BF16, dev.index)
        handle = C.c_void_p()
        with torch.cuda.device(dev):
            L.check(self.lib.sf_plan_create(C.byref(geo), C.byref(handle)), "sf_plan_create")
        self.plan = handle
        self._keep: List[torch.Tensor] = []          # packed weights / vectors referenced by the plan
        self.act: Dict[int, Tuple[torch.Tensor, Optional[torch.Tensor]]] = {}
        self._alloc_workspace(path_slots)
        self.load_weights(sd, prefix)
        self.n_obs_images = 0
        self.launches = 0

    # ------------------------------------------------------------------ workspace
    def _new_act(self, buf, n_images, channels):
        self.alloc_gen = getattr(self, "alloc_gen", 0) + 1     # captured CUDA graphs hold buffer addresses: (re)allocation invalidates them
        shape = (n_images, self.H, self.W, channels)
        hi = torch.zeros(shape, dtype=torch.bfloat16, device=self.device)
        lo = torch.zeros(shape, dtype=torch.bfloat16, device=self.device) if self.x3 else None
        self.act[buf] = (hi, lo)
        L.check(self.lib.sf_plan_bind_act(self.plan, buf, hi.data_ptr(), lo.data_ptr() if lo is not None else None, channels,
                                          n_images), "sf_plan_bind_act")

    def _bind_f32(self, slot, t):
        L.check(self.lib.sf_plan_bind_f32(self.plan, slot, t.data_ptr() if t is not None else None), "sf_plan_bind_f32")

    def _alloc_workspace(self, path_slots):
        B, H, W, Cc = self.max_images, self.H, self.W, self.C
        for buf, mult in _BUF_CHANNELS.items():
            if self.se_fold and buf in (BUF_Y1, BUF_Y2):
                continue                              # SE outputs are never materialised when the layers are folded
            self._new_act(buf, B, Cc * mult)
        self._new_act(BUF_ZERO, 1, Cc)
        f32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)
        self.state32 = [f32(B, H, W, Cc), f32(B, H, W, Cc)]
        self.a32, self.b32 = f32(B, H, W, Cc), f32(B, H, W, Cc)
        # SE scratch: partial channel sums [2][B][SE_MAX_PARTIALS][2C] | scales [2][B][2C] | block counters [2][B] (zero-initialised)
        P = L.SE_MAX_PARTIALS
        self.se_scratch = f32(2 * B * P * 2 * Cc + 2 * B * 2 * Cc + 2 * B)
        self.se_sums = self.se_scratch[: 2 * B * P * 2 * Cc].view(2, B, P, 2 * Cc)
        self.x32, self.params32 = f32(B, H, W, Cc), f32(B, H, W, 2 * Cc)
        self.errflag = torch.zeros(1, dtype=torch.int32, device=self.device)
        for slot, t in ((L.F32_STATE0, self.state32[0]), (L.F32_STATE1, self.state32[1]), (L.F32_A, self.a32), (L.F32_B, self.b32),
                        (L.F32_SE_SUMS, self.se_scratch), (L.F32_X, self.x32), (L.F32_PARAMS, self.params32),
                        (L.F32_ERRFLAG, self.errflag)):
            self._bind_f32(slot, t)
        self.path = None
        self.ensure_path_slots(max(1, path_slots))
        self.eps = None

    def ensure_path_slots(self, n):
        if self.path is None or self.path.shape[0] < n:
            self.alloc_gen = getattr(self, "alloc_gen", 0) + 1
            self.path = torch.zeros((n, self.H, self.W, self.C), dtype=torch.float32, device=self.device)
            self._bind_f32(L.F32_PATH, self.path)

    def bind_eps(self, eps: torch.Tensor):
        """eps: fp32 NCHW [slots, 64, H, W] standard-normal noise (torch's own generation order)."""
        assert eps.dtype == torch.float32 and eps.is_contiguous() and tuple(eps.shape[1:]) == (self.C, self.H, self.W), eps.shape
        self.eps = eps
        self._bind_f32(L.F32_EPS, eps)

    # ------------------------------------------------------------------ weights
    def load_weights(self, sd: Dict[str, torch.Tensor], prefix: str):
        """(Re)packs every stage from reference-named parameters ``{prefix}gru_c.*``, ``{prefix}gru_obs.gru_d.*``,
        ``{prefix}p_model.*``."""
        from . import cpack

        sd = {k: v.detach() for k, v in sd.items() if k.startswith(prefix)}
        pre = prefix
        # captured CUDA graphs hold the addresses of the packed weights / vectors released below: retire them
        self.alloc_gen = getattr(self, "alloc_gen", 0) + 1
        self._keep = []
        self.stage_defs: Dict[int, object] = {}          # slot -> cpack.PackedItem (name, epilogue, chunks, flags, ...)
        self.stage_names: Dict[int, str] = {}
        pair = os.environ.get("SF_PAIR_ROWS", "1") != "0"
        b2b = os.environ.get("SF_B2B", "1") != "0"
        pair3 = os.environ.get("SF_PAIR_3X3", "0") == "1"        # measured slower (decode 38 -> 45 us): off unless asked for
        # BatchNorm / cat[state, state] folding, tap order and the hi / lo split happen in the library (sf_pack_cell_weights,
        # sf_pack_pmodel_weights: csrc/sf_ode.cu) -- the same bytes a C host gets; cell_stage_defs / prior_stage_defs / pack_stage
        # above restate them in torch and are pinned to the library's output by tests/test_c_host.py
        cells = [cpack.pack_cell(sd, pre + "gru_c.", self.x3, pair, b2b, pair3), cpack.pack_cell(sd, pre + "gru_obs.gru_d.", self.x3, pair, b2b, pair3)]
        n_cell = len(cells[0])
        cell_slots = [list(range(ws * n_cell, (ws + 1) * n_cell)) for ws in range(2)]
        for ws in range(2):
            for slot, item in zip(cell_slots[ws], cells[ws]):
                self.stage_defs[slot] = item
                if ws == 0:
                    self.stage_names[slot] = item.name
        prior_items, slot, se_items = [], 2 * n_cell, []
        for it in cpack.pack_pmodel(sd, pre + "p_model.", self.x3, self.se_fold, pair3):
            if it.se_layer >= 0:
                prior_items.append((L.SE_FOLD_ITEM_BASE if self.se_fold else L.SE_ITEM_BASE) + it.se_layer)
                self.stage_names[prior_items[-1]] = "se" + str(it.se_layer + 1)
                se_items.append(it)
            else:
                self.stage_defs[slot] = it
                self.stage_names[slot] = it.name
                prior_items.append(slot)
                slot += 1
        self.cell_slots, self.prior_items = cell_slots, prior_items
        for slot, item in self.stage_defs.items():
            wp = item.w.to(self.device)
            vec = item.vec.to(self.device, torch.float32).contiguous()
            arr = (L.Chunk * len(item.chunks))(*[L.Chunk(**c) for c in item.chunks])
            io = (C.c_int32 * max(1, len(item.io)))(*item.io)
            io_off = (C.c_int32 * max(1, len(item.io)))(*item.io_off)
            self._keep += [wp, vec]
            L.check(self.lib.sf_plan_define_stage(self.plan, slot, item.epilogue, len(item.chunks), arr, wp.data_ptr(), wp.shape[0],
                                                  vec.data_ptr(), vec.numel(), io, io_off, len(item.io), item.flags),
                    f"sf_plan_define_stage({slot}:{item.name})")
            if item.fold_se is not None:
                w32, meta = item.w32.to(self.device), item.row_meta.to(self.device)
                scaled = torch.zeros((self.max_images, wp.shape[0], 64), dtype=torch.bfloat16, device=self.device)
                self._keep += [w32, meta, scaled]
                L.check(self.lib.sf_plan_define_stage_fold(self.plan, slot, item.fold_se, w32.data_ptr(), meta.data_ptr(), scaled.data_ptr()),
                        f"sf_plan_define_stage_fold({slot}:{item.name})")
        for it, (zin, yout) in zip(se_items, ((BUF_Z1, BUF_Y1), (BUF_Z2, BUF_Y2))):
            fc1, fc2 = it.fc1.to(self.device), it.fc2.to(self.device)
            self._keep += [fc1, fc2]
            L.check(self.lib.sf_plan_define_se(self.plan, it.se_layer, fc1.data_ptr(), fc2.data_ptr(), zin, yout), "sf_plan_define_se")
        i32 = lambda v: (C.c_int32 * len(v))(*v)
        L.check(self.lib.sf_plan_define_event_graph(self.plan, i32(cell_slots[0]), i32(cell_slots[1]), n_cell,
                                                    i32(prior_items), len(prior_items)), "sf_plan_define_event_graph")
        with torch.cuda.device(self.device):
            L.check(self.lib.sf_plan_finalize(self.plan), "sf_plan_finalize")

    # ------------------------------------------------------------------ data movement
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def pack_into(self, buf: int, src_nchw: torch.Tensor, img_offset: int = 0):
        """NCHW fp32 -> the NHWC bf16 activation buffer ``buf`` (images [img_offset, img_offset + n))."""
        src = src_nchw.contiguous().float()
        n, c, h, w = src.shape
        assert (h, w) == (self.H, self.W)
        hi, lo = self.act[buf]
        assert img_offset + n <= hi.shape[0] and c == hi.shape[3]
        L.check(self.lib.sf_pack_nchw_f32(src.data_ptr(), hi[img_offset:].data_ptr(), lo[img_offset:].data_ptr() if lo is not None else None,
                                          n, c, h, w, self._stream()), "sf_pack_nchw_f32")

    def reserve_observations(self, n_images: int):
        """An engine-owned observation buffer of at least n_images (re-bound if the slot currently points at somebody else's
        planes, e.g. the fused encoder's output: a captured graph must never keep reading a buffer the engine does not own)."""
        if BUF_OBS not in self.act or not getattr(self, "_obs_owned", False) or self.act[BUF_OBS][0].shape[0] < n_images:
            self._new_act(BUF_OBS, n_images, self.C)
            self._obs_owned = True
        self.n_obs_images = n_images

    def bind_observations(self, hx_nchw: torch.Tensor):
        """Encoded observations [n_img, 64, H, W] fp32 -> OBS activation buffer (the jump cell's x input)."""
        n = hx_nchw.shape[0]
        self.reserve_observations(n)
        self.pack_into(BUF_OBS, hx_nchw)

    def bind_observation_planes(self, hi: torch.Tensor, lo: Optional[torch.Tensor]):
        """Use already-encoded NHWC bf16 planes [n, H, W, C] (the fused encoder's output buffer) as the observation buffer."""
        assert tuple(hi.shape[1:]) == (self.H, self.W, self.C) and hi.dtype == torch.bfloat16 and hi.is_contiguous()
        if self.x3 and lo is None:
            raise L.SfError("bf16x3 needs the residual plane of the observations")
        cur = self.act.get(BUF_OBS)
        same = cur is not None and cur[0].data_ptr() == hi.data_ptr() and cur[0].shape[0] == hi.shape[0] and \
            (cur[1] is None) == (lo is None) and (lo is None or cur[1].data_ptr() == lo.data_ptr())
        if not same:
            self.alloc_gen = getattr(self, "alloc_gen", 0) + 1      # the OBS slot's address changes: graphs captured with the old one retire
            self.act[BUF_OBS] = (hi, lo)
            self._obs_owned = False
            L.check(self.lib.sf_plan_bind_act(self.plan, BUF_OBS, hi.data_ptr(), lo.data_ptr() if lo is not None else None, self.C, hi.shape[0]),
                    "sf_plan_bind_act")
        self.n_obs_images = hi.shape[0]

    def set_state(self, which: int, state_nchw: torch.Tensor):
        n = state_nchw.shape[0]
        nhwc = state_nchw.float().permute(0, 2, 3, 1).contiguous()
        self.state32[which][:n].copy_(nhwc)
        self.pack_into(BUF_S0 + which, state_nchw)

    def zero_state(self, which: int = 0):
        self.state32[which].zero_()
        hi, lo = self.act[BUF_S0 + which]
        hi.zero_()
        if lo is not None:
            lo.zero_()

    def snapshot(self, n: int):
        """Copies of everything an event reads from a previous one for the first n samples (state master + operand planes, the
        sampled input's operand planes): StreamingOdeSession rolls a prediction forward and then restores them."""
        keep = [self.state32[0][:n].clone()]
        for buf in (BUF_S0, BUF_X):
            keep += [None if p is None else p[:n].clone() for p in self.act[buf]]
        return keep

    def restore(self, snap, n: int):
        self.state32[0][:n].copy_(snap[0])
        k = 1
        for buf in (BUF_S0, BUF_X):
            for p in self.act[buf]:
                if p is not None:
                    p[:n].copy_(snap[k])
                k += 1

    def unpack_path(self, slots, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Recorded states (NHWC fp32 path buffer) -> NCHW fp32 [len(slots), 64, H, W]."""
        idx = slots if isinstance(slots, torch.Tensor) else torch.tensor(list(slots), dtype=torch.int32, device=self.device)
        n = idx.numel()
        out = torch.empty((n, self.C, self.H, self.W), dtype=torch.float32, device=self.device) if out is None else out
        L.check(self.lib.sf_unpack_nhwc_f32(self.path.data_ptr(), out.data_ptr(), idx.data_ptr(), n, self.C, self.H, self.W,
                                            self._stream()), "sf_unpack_nhwc_f32")
        return out

    def unpack_f32(self, t_nhwc: torch.Tensor, n: int) -> torch.Tensor:
        c = t_nhwc.shape[3]
        out = torch.empty((n, c, self.H, self.W), dtype=torch.float32, device=self.device)
        L.check(self.lib.sf_unpack_nhwc_f32(t_nhwc.data_ptr(), out.data_ptr(), None, n, c, self.H, self.W, self._stream()),
                "sf_unpack_nhwc_f32")
        return out

    # ------------------------------------------------------------------ events
    @staticmethod
    def build_table(events: List[dict]) -> Tuple[np.ndarray, List[L.Event]]:
        """events: dicts with kind, samples, x_img, rec, eps, dt (lists of equal length) + x_buf, s_in, s_base, s_out,
        run_cell, run_prior, want_f32.  Returns the packed int32 table and the sf_event structs."""
        rows, evs, off = [], [], 0
        for e in events:
            n = len(e["samples"])
            blk = np.empty((5, n), dtype=np.int32)
            blk[0] = e["samples"]
            blk[1] = e["x_img"]
            blk[2] = e.get("rec", [-1] * n)
            blk[3] = e.get("eps", [0] * n)
            blk[4] = np.asarray(e.get("dt", [0.0] * n), dtype=np.float64).astype(np.float32).view(np.int32)
            rows.append(blk.reshape(-1))
            evs.append(L.Event(e["kind"], n, e["x_buf"], e.get("s_in", 0), e.get("s_base", 0), e.get("s_out", 0),
                               int(e.get("run_cell", 1)), int(e.get("run_prior", 1)), int(e.get("want_f32", 0)), off))
            off += 5 * n
        return (np.concatenate(rows) if rows else np.zeros(0, np.int32)), evs

    def upload_table(self, table: np.ndarray) -> torch.Tensor:
        t = torch.from_numpy(table)
        return t.to(self.device, non_blocking=False)

    def run_events(self, evs: List[L.Event], table_dev: torch.Tensor):
        arr = (L.Event * len(evs))(*evs)
        with torch.cuda.device(self.device):
            L.check(self.lib.sf_plan_run_events(self.plan, arr, len(evs), table_dev.data_ptr(), self._stream()), "sf_plan_run_events")
        n = self.lib.sf_plan_last_launches(self.plan)
        self.launches += n
        return n

    def run_rollout(self, events: List[dict]) -> int:
        """events (dicts, see build_table) -> one table upload + one C call that enqueues every stage launch."""
        table, evs = self.build_table(events)
        return self.run_events(evs, self.upload_table(table))

    def run_stage(self, stage: int, ev: L.Event, table_dev: torch.Tensor):
        with torch.cuda.device(self.device):
            L.check(self.lib.sf_plan_run_stage(self.plan, stage, C.byref(ev), table_dev.data_ptr(), self._stream()), "sf_plan_run_stage")

    def check_errflag(self):
        v = int(self.errflag.item())
        if v:
            raise L.SfError(f"device-side pipeline timeout, code 0x{v:x}")

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.lib.sf_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass
