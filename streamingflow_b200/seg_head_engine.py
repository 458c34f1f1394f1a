"""The BEV Decoder head on the conv-stage kernels ("next" row 3 of SURVEY.md 8f): streamingflow/models/decoder.py:91-140.

    x [n, 64, H, W] -> first_conv 7x7/2 + BN + ReLU -> layer1 (2 BasicBlocks, 64 @ H/2) -> layer2 (64 -> 128, /2) -> layer3
    (128 -> 256, /2) -> UpsamplingAdd x3 (bilinear x2, 1x1 conv, BN, + skip) -> heads (3x3 conv + BN + ReLU, 1x1 conv -> K)

Every convolution runs on the tcgen05 implicit-GEMM stage kernel of the ODE loop:
  * stride-2 convolutions: the input is regrouped once into its four pixel phases (sf_space_to_depth2: channel block 2 py + px
    of the half-resolution buffer holds x[2i + py, 2j + px]); the strided conv is then a stride-1 conv over the phase images --
    a 3x3 chunk for the taps that form a full 3x3 block of a phase, 1x1 chunks with shifted windows for the rest (the mechanism
    of the dilated ASPP taps).  Out-of-image taps are TMA zero fill, exactly the reference's zero padding;
  * BasicBlock: conv1 = bias_act(ReLU), conv2 = res_id with the activation applied after the residual add; a block with a
    down-sampling shortcut accumulates conv2 AND the strided 1x1 shortcut into the same accumulator (both are linear, BatchNorms
    folded), bias = b2 + b_shortcut, ReLU;
  * UpsamplingAdd: the 1x1 conv + BN commute with the bilinear interpolation (its weights sum to one), so they run at the LOW
    resolution (4x fewer MACs) and one elementwise kernel does bilinear x2 + skip add (sf_bilinear_up2_add);
  * heads: 3x3 conv + BN + ReLU as a stage, the K <= 4 output channels (+ arg-max for the segmentation masks) by sf_head_1x1.
Batch = all n = b*s frames at once; one sf_plan per resolution level.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib as L
from .codec_engine import LevelPlan, _act_flags
from .engine import StageDef


def _fold_bn(w: torch.Tensor, sd, bn: str):
    scale = sd[bn + ".weight"].float() / torch.sqrt(sd[bn + ".running_var"].float() + 1e-5)
    return w.float() * scale[:, None, None, None], sd[bn + ".bias"].float() - sd[bn + ".running_mean"].float() * scale


def add_strided(st: StageDef, phase_buf: int, w: torch.Tensor, pad: int, col: int, init: bool) -> bool:
    """Adds the chunks of a stride-2 convolution with filter w [n, cin, R, R] and padding ``pad`` over the space-to-depth
    buffer ``phase_buf`` (channel block 2 py + px = phase (py, px), cin channels each).  Output (y, x) reads input
    (2y + dy - pad, 2x + dx - pad) = phase (py, px) pixel (y + a, x + b) with dy - pad = 2a + py, dx - pad = 2b + px.
    Returns the ``init`` flag for the next add (False once something was added)."""
    n, cin, R, _ = w.shape
    first = bool(init)
    for py in range(2):
        for px in range(2):
            taps_y = [(dy, (dy - pad - py) // 2) for dy in range(R) if (dy - pad - py) % 2 == 0]
            taps_x = [(dx, (dx - pad - px) // 2) for dx in range(R) if (dx - pad - px) % 2 == 0]
            c0 = (2 * py + px) * cin
            ay, ax = {a: dy for dy, a in taps_y}, {b: dx for dx, b in taps_x}
            block = all(a in ay for a in (-1, 0, 1)) and all(b in ax for b in (-1, 0, 1))
            if block:          # the taps a, b in {-1, 0, 1} of this phase as ONE 3x3 chunk
                w3 = torch.stack([torch.stack([w[:, :, ay[a], ax[b]] for b in (-1, 0, 1)], dim=-1) for a in (-1, 0, 1)], dim=-2)
                st.add(phase_buf, w3.contiguous(), col, first, c0=c0)
                first = False
            for dy, a in taps_y:
                for dx, b in taps_x:
                    if block and a in (-1, 0, 1) and b in (-1, 0, 1):
                        continue
                    st.add(phase_buf, w[:, :, dy:dy + 1, dx:dx + 1].contiguous(), col, first, c0=c0, ox=b, oy=a)
                    first = False
    return first


# buffer ids per level
(X, U1, HD) = range(3)                                  # level 0: input frames, up1 output, head hidden
(P0, F, T1, A1, S2, U2, V1) = range(7)                  # level 1
(P1, T2, A2, S3, U3, V2) = range(6)                     # level 2
(P2, T3, A3, X3, V3) = range(5)                         # level 3
LEVEL_BUFS = [{X: 64, U1: 64, HD: 64}, {P0: 256, F: 64, T1: 64, A1: 64, S2: 64, U2: 64, V1: 64},
              {P1: 256, T2: 128, A2: 128, S3: 128, U3: 128, V2: 64}, {P2: 512, T3: 256, A3: 256, X3: 256, V3: 128}]
HEADS = (("segmentation", "segmentation_head", False), ("pedestrian", "pedestrian_head", False), ("hdmap", "hdmap_head", False),
         ("instance_center", "instance_center_head", True), ("instance_offset", "instance_offset_head", False),
         ("instance_flow", "instance_future_head", False), ("costvolume", "costvolume_head", False))


def _basic_block(name, sd, p, src, tmp, dst, cin, cout, phase_buf=None):
    """torchvision BasicBlock (eval): relu(bn2(conv2(relu(bn1(conv1(x))))) + shortcut(x)); stride 2 + 1x1 shortcut when
    phase_buf is given (x then lives in the space-to-depth buffer).  Launches of at most 128 output channels."""
    w1, b1 = _fold_bn(sd[p + ".conv1.weight"], sd, p + ".bn1")
    w2, b2 = _fold_bn(sd[p + ".conv2.weight"], sd, p + ".bn2")
    stages = []
    nh = max(1, cout // 128)
    n = cout // nh
    for h in range(nh):
        r = slice(n * h, n * h + n)
        st = StageDef(f"{name}.c1{'ab'[h] if nh > 1 else ''}", L.EPI_BIAS_LRELU, b1[r], [tmp], [n * h], flags=_act_flags(L.ACT_RELU))
        if phase_buf is None:
            st.add(src, w1[r], 0, 1)
        else:
            add_strided(st, phase_buf, w1[r], 1, 0, True)
        stages.append(st)
    if phase_buf is None:
        for h in range(nh):
            r = slice(n * h, n * h + n)
            stages.append(StageDef(f"{name}.c2{'ab'[h] if nh > 1 else ''}", L.EPI_RES_ID, b2[r], [src, dst], [n * h, n * h],
                                   flags=_act_flags(L.ACT_RELU) | L.FLAG_ACT_AFTER_RES).add(tmp, w2[r], 0, 1))
    else:
        wd, bd = _fold_bn(sd[p + ".downsample.0.weight"], sd, p + ".downsample.1")
        for h in range(nh):
            r = slice(n * h, n * h + n)
            st = StageDef(f"{name}.c2{'ab'[h] if nh > 1 else ''}", L.EPI_BIAS_LRELU, (b2 + bd)[r], [dst], [n * h], flags=_act_flags(L.ACT_RELU))
            st.add(tmp, w2[r], 0, 1)
            add_strided(st, phase_buf, wd[r], 0, 0, False)
            stages.append(st)
    return stages


def _conv1x1_bn(name, sd, p, src, dst):
    w, b = _fold_bn(sd[p + ".upsample_layer.1.weight"], sd, p + ".upsample_layer.2")
    return StageDef(name, L.EPI_BIAS_LRELU, b, [dst], flags=_act_flags(L.ACT_NONE)).add(src, w, 0, 1)


def decoder_graph(sd: Dict[str, torch.Tensor], heads: List[str]):
    """Ops of Decoder.forward over the four levels: ("stage", level, StageDef) | ("s2d", (level, buf), (level, buf), C) |
    ("upadd", (level, buf) low-res source, (level, buf) skip, (level, buf) dst, C) | ("head", name, K, sigmoid)."""
    g = [("s2d", (0, X), (1, P0), 64)]
    w, b = _fold_bn(sd["first_conv.weight"], sd, "bn1")
    st = StageDef("first_conv", L.EPI_BIAS_LRELU, b, [F], flags=_act_flags(L.ACT_RELU))
    add_strided(st, P0, w, 3, 0, True)
    g.append(("stage", 1, st))
    g += [("stage", 1, s) for s in _basic_block("l1.0", sd, "layer1.0", F, T1, A1, 64, 64) + _basic_block("l1.1", sd, "layer1.1", A1, T1, S2, 64, 64)]
    g.append(("s2d", (1, S2), (2, P1), 64))
    g += [("stage", 2, s) for s in _basic_block("l2.0", sd, "layer2.0", None, T2, A2, 64, 128, phase_buf=P1)
          + _basic_block("l2.1", sd, "layer2.1", A2, T2, S3, 128, 128)]
    g.append(("s2d", (2, S3), (3, P2), 128))
    g += [("stage", 3, s) for s in _basic_block("l3.0", sd, "layer3.0", None, T3, A3, 128, 256, phase_buf=P2)
          + _basic_block("l3.1", sd, "layer3.1", A3, T3, X3, 256, 256)]
    g.append(("stage", 3, _conv1x1_bn("up3.conv", sd, "up3_skip", X3, V3)))
    g.append(("upadd", (3, V3), (2, S3), (2, U3), 128))
    g.append(("stage", 2, _conv1x1_bn("up2.conv", sd, "up2_skip", U3, V2)))
    g.append(("upadd", (2, V2), (1, S2), (1, U2), 64))
    g.append(("stage", 1, _conv1x1_bn("up1.conv", sd, "up1_skip", U2, V1)))
    g.append(("upadd", (1, V1), (0, X), (0, U1), 64))
    for key, mod, sig in HEADS:
        if key in heads:
            w, b = _fold_bn(sd[mod + ".0.weight"], sd, mod + ".1")
            g.append(("stage", 0, StageDef(mod + ".conv", L.EPI_BIAS_LRELU, b, [HD], flags=_act_flags(L.ACT_RELU)).add(U1, w, 0, 1)))
            g.append(("head", key, mod, sig))
    return g


class SegHeadEngine:
    """Decoder.forward for n frames of a fixed BEV size (64 input channels; H, W multiples of 8)."""

    def __init__(self, sd: Dict[str, torch.Tensor], H: int, W: int, n: int, precision: str, device, heads: List[str]):
        if H % 8 or W % 8:
            raise L.SfError("the fused Decoder head needs BEV height / width that are multiples of 8 (three stride-2 stages)")
        self.lib = L.load()
        self.H, self.W, self.n, self.device = H, W, n, device
        self.x3 = precision == "bf16x3"
        sd = {k: v.detach().to(device) for k, v in sd.items()}
        if sd["first_conv.weight"].shape[:2] != (64, 64):
            raise L.SfError("the fused Decoder head is built for 64 input channels")
        self.dims = [(H >> i, W >> i) for i in range(4)]
        self.plans = [LevelPlan(self.lib, h, w, n, self.x3, device) for h, w in self.dims]
        for lvl, bufs in enumerate(LEVEL_BUFS):
            for b, ch in bufs.items():
                if not (lvl == 0 and b == X):          # the input planes are bound per call
                    self.plans[lvl].buf(b, ch)
        self.own_x = None
        self.ops, self.head_params = [], {}
        for op in decoder_graph(sd, heads):
            if op[0] == "stage":
                slot = self.plans[op[1]].stage(op[2])
                if self.ops and self.ops[-1][0] == "stages" and self.ops[-1][1] == op[1]:
                    self.ops[-1][2].append(slot)
                else:
                    self.ops.append(("stages", op[1], [slot]))
            elif op[0] == "head":
                _, key, mod, sig = op
                w = sd[mod + ".3.weight"].float()[:, :, 0, 0].contiguous()
                self.head_params[key] = (w, sd[mod + ".3.bias"].float().contiguous(), int(w.shape[0]), sig)
                self.ops.append(op)
            else:
                self.ops.append(op)
        for p in self.plans:
            p.finalize()
        self.launches = 0

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _bind_x(self, planes, n):
        if planes is None:
            if self.own_x is None:
                shape = (self.n, self.H, self.W, 64)
                self.own_x = (torch.zeros(shape, dtype=torch.bfloat16, device=self.device),
                              torch.zeros(shape, dtype=torch.bfloat16, device=self.device) if self.x3 else None)
            planes = self.own_x
        self.plans[0].buf(X, 64, planes=planes)
        return planes

    def run(self, x: Optional[torch.Tensor] = None, planes=None, frames: Optional[List[int]] = None, want_mask: bool = True):
        """x: [n, 64, H, W] fp32 NCHW frames -- or ``planes`` = (hi, lo) NHWC bf16 [n, H, W, 64] already in engine layout (the
        fused refinement's output buffer).  Returns {head: fp32 NCHW [n, K, H, W]} and, for 'segmentation' with want_mask, the
        uint8 arg-max masks [n, H, W] under 'segmentation_argmax'."""
        lib, stream = self.lib, self._stream()
        if planes is None:
            n = x.shape[0]
            assert n <= self.n and tuple(x.shape[1:]) == (64, self.H, self.W)
            planes = self._bind_x(None, n)
            src = x.contiguous().float()
            L.check(lib.sf_pack_nchw_f32(src.data_ptr(), planes[0].data_ptr(), planes[1].data_ptr() if planes[1] is not None else None, n, 64,
                                         self.H, self.W, stream), "pack")
            self.launches += 1
        else:
            n = planes[0].shape[0]
            assert n <= self.n and tuple(planes[0].shape[1:]) == (self.H, self.W, 64)
            if self.x3 and planes[1] is None:
                raise L.SfError("bf16x3 needs the residual plane of the input frames")
            self._bind_x((planes[0], planes[1] if self.x3 else None), n)
        out = {}
        ptr = lambda t: t.data_ptr() if t is not None else None
        for op in self.ops:
            if op[0] == "stages":
                self.launches += self.plans[op[1]].run(op[2], n)
            elif op[0] == "s2d":
                (ls, bs), (ld, bd), ch = op[1], op[2], op[3]
                h, w = self.dims[ls]
                for s_, d_ in zip(self.plans[ls].bufs[bs], self.plans[ld].bufs[bd]):
                    if s_ is not None and d_ is not None:
                        L.check(lib.sf_space_to_depth2(s_.data_ptr(), d_.data_ptr(), n, h, w, ch, stream), "space_to_depth2")
                        self.launches += 1
            elif op[0] == "upadd":
                (ls, bs), (lk, bk), (ld, bd), ch = op[1], op[2], op[3], op[4]
                h, w = self.dims[ls]
                s_, k_, d_ = self.plans[ls].bufs[bs], self.plans[lk].bufs[bk], self.plans[ld].bufs[bd]
                L.check(lib.sf_bilinear_up2_add(ptr(s_[0]), ptr(s_[1]), ptr(k_[0]), ptr(k_[1]), ptr(d_[0]), ptr(d_[1]), n, h, w, ch, stream),
                        "bilinear_up2_add")
                self.launches += 1
            else:
                _, key, mod, sig = op
                w, b, K, _ = self.head_params[key]
                hd = self.plans[0].bufs[HD]
                res = torch.empty((n, K, self.H, self.W), dtype=torch.float32, device=self.device)
                mask = torch.empty((n, self.H, self.W), dtype=torch.uint8, device=self.device) if (key == "segmentation" and want_mask) else None
                L.check(lib.sf_head_1x1(ptr(hd[0]), ptr(hd[1]), w.data_ptr(), b.data_ptr(), K, int(sig), res.data_ptr(), ptr(mask), n, self.H, self.W,
                                        stream), "head_1x1")
                self.launches += 1
                out[key] = res
                if mask is not None:
                    out["segmentation_argmax"] = mask
        return out
