"""SmallEncoder / SmallDecoder on the conv-stage kernels ("next" row 1 of SURVEY.md 8f).

The reference's encoder (res_models.py:82-109) and decoder (:112-147) are chains of 3x3 convolutions with BatchNorm +
LeakyReLU, residual adds / 1x1 projections, two 2x2 max-pools and two nearest x2 up-samplings.  Every convolution maps onto
the same tcgen05 implicit-GEMM stage kernel the ODE loop uses (epilogues bias_act / res_id / res_proj, BatchNorm folded,
ConvTranspose2d(stride 1) rewritten as a convolution with the flipped, transposed filter); pool / up-sample / casts are small
NHWC kernels.  One sf_plan per resolution level.  The encoder's output IS the ODE engine's observation buffer (NHWC bf16
planes) and the decoder reads the recorded path states directly, so no NCHW fp32 round trip remains between the three parts.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib as L
from .engine import StageDef, _bn_fold, pack_stage, pair_rows_if_eligible


def _convT_as_conv(w_t: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d(stride 1, padding 1, k 3) weight [in, out, k, k] -> the equivalent Conv2d weight [out, in, k, k]."""
    return w_t.permute(1, 0, 2, 3).flip(2, 3).contiguous()


def _bn_fold_w(w: torch.Tensor, sd, norm: str):
    scale = sd[norm + ".weight"].float() / torch.sqrt(sd[norm + ".running_var"].float() + 1e-5)
    return w * scale[:, None, None, None], sd[norm + ".bias"].float() - sd[norm + ".running_mean"].float() * scale


class LevelPlan:
    """One sf_plan at a fixed resolution with its own NHWC activation buffers and conv stages."""

    def __init__(self, lib, H, W, max_images, x3, device, C_hidden: int = 64):
        """C_hidden: the hidden width the gates / propose epilogues are instantiated for (64 or 128); the bias_act / residual
        epilogues are width-generic (<= 128 output channels per launch)."""
        self.lib, self.H, self.W, self.n, self.x3, self.device = lib, H, W, max_images, x3, device
        self.C_hidden = C_hidden
        geo = L.Geometry(max_images, H, W, C_hidden, L.PREC_BF16X3 if x3 else L.PREC_BF16, device.index)
        h = C.c_void_p()
        with torch.cuda.device(device):
            L.check(lib.sf_plan_create(C.byref(geo), C.byref(h)), "sf_plan_create")
        self.plan = h
        self.bufs: Dict[int, Tuple[torch.Tensor, Optional[torch.Tensor]]] = {}
        self.keep: List[torch.Tensor] = []
        self.errflag = torch.zeros(1, dtype=torch.int32, device=device)
        L.check(lib.sf_plan_bind_f32(self.plan, L.F32_ERRFLAG, self.errflag.data_ptr()), "bind errflag")
        self.n_stages = 0
        self._tables: Dict[int, Tuple[torch.Tensor, L.Event]] = {}

    def buf(self, buf_id: int, channels: int, planes=None):
        if planes is None:
            shape = (self.n, self.H, self.W, channels)
            hi = torch.zeros(shape, dtype=torch.bfloat16, device=self.device)
            lo = torch.zeros(shape, dtype=torch.bfloat16, device=self.device) if self.x3 else None
        else:
            hi, lo = planes
        self.bufs[buf_id] = (hi, lo)
        L.check(self.lib.sf_plan_bind_act(self.plan, buf_id, hi.data_ptr(), lo.data_ptr() if lo is not None else None, channels, hi.shape[0]),
                "sf_plan_bind_act")
        return hi, lo

    def bind_out32(self, t: torch.Tensor):
        self.out32 = t
        L.check(self.lib.sf_plan_bind_f32(self.plan, L.F32_OUT, t.data_ptr()), "bind out32")

    def stage(self, sdef: StageDef) -> int:
        slot = self.n_stages
        self.n_stages += 1
        pair_rows_if_eligible(sdef, self.C_hidden)
        chunks, wp = pack_stage(sdef, self.x3)
        vec = sdef.vec.to(torch.float32).contiguous()
        arr = (L.Chunk * len(chunks))(*[L.Chunk(**c) for c in chunks])
        io = (C.c_int32 * max(1, len(sdef.io)))(*sdef.io)
        io_off = (C.c_int32 * max(1, len(sdef.io)))(*sdef.io_off)
        self.keep += [wp, vec]
        L.check(self.lib.sf_plan_define_stage(self.plan, slot, sdef.epilogue, len(chunks), arr, wp.data_ptr(), wp.shape[0], vec.data_ptr(),
                                              vec.numel(), io, io_off, len(sdef.io), sdef.flags), f"define_stage({sdef.name})")
        return slot

    def finalize(self):
        with torch.cuda.device(self.device):
            L.check(self.lib.sf_plan_finalize(self.plan), "sf_plan_finalize")

    def run(self, slots, n_images: int):
        if n_images not in self._tables:
            rows = np.zeros((5, n_images), dtype=np.int32)
            rows[0] = np.arange(n_images)
            rows[1] = np.arange(n_images)
            rows[2] = -1
            self._tables[n_images] = (torch.from_numpy(rows.reshape(-1)).to(self.device), L.Event(0, n_images, 0, 0, 0, 0, 1, 0, 0, 0))
        tdev, ev = self._tables[n_images]
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            for s in slots:
                L.check(self.lib.sf_plan_run_stage(self.plan, s, C.byref(ev), tdev.data_ptr(), stream), "sf_plan_run_stage")
        return len(slots)

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.lib.sf_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass


def _act_flags(act: int, out32: bool = False) -> int:
    return (act << 1) | (L.FLAG_OUT32 if out32 else 0)


def _conv_bn_act(name, sd, p, src, dst, act=L.ACT_LRELU, w=None):
    """ConvBlock (conv3x3 + BN + activation) as launches of 128 (or fewer) output channels."""
    if w is None:
        w = sd[p + ".conv.weight"].float()
    w, b = _bn_fold_w(w, sd, p + ".norm")
    n = min(w.shape[0], 128)
    out = []
    for h in range(w.shape[0] // n):
        r = slice(n * h, n * h + n)
        out.append(StageDef(f"{name}{'abcd'[h] if w.shape[0] > n else ''}", L.EPI_BIAS_LRELU, b[r], [dst], [n * h], flags=_act_flags(act)).add(src, w[r], 0, 1))
    return out


def _res_block(name, sd, p, src, tmp, dst, cin, cout):
    """ResBlock(cin -> cout): conv_1 (cin -> cin), conv_2 (cin -> cout), identity or 1x1-projected skip (res_models.py:52-79);
    every launch owns <= 128 output channels."""
    stages = _conv_bn_act(name + ".c1", sd, p + ".layers.conv_1", src, tmp)
    w2, b2 = _bn_fold_w(sd[p + ".layers.conv_2.conv.weight"].float(), sd, p + ".layers.conv_2.norm")
    n = min(cout, 128)
    for h in range(cout // n):
        r, tag = slice(n * h, n * h + n), ('abcd'[h] if cout > n else '')
        if cin == cout:
            stages.append(StageDef(f"{name}.c2{tag}", L.EPI_RES_ID, b2[r], [src, dst], [n * h, n * h]).add(tmp, w2[r], 0, 1))
        else:
            wp, bp = sd[p + ".projection.weight"].float(), sd[p + ".projection.bias"].float()
            stages.append(StageDef(f"{name}.c2{tag}", L.EPI_RES_PROJ, torch.cat([b2[r], bp[r]]), [dst], [n * h])
                          .add(tmp, w2[r], 0, 1).add(src, wp[r], n, 1))
    return stages


def codec_widths(sd, e="srvp_encoder."):
    """(nc, nf): input = latent width (MODEL.ENCODER.OUT_CHANNELS) and filter size (MODEL.SMALL_ENCODER.FILTER_SIZE)."""
    return sd[e + "blocks.0.layers.conv_1.conv.weight"].shape[0], sd[e + "blocks.0.layers.conv_2.conv.weight"].shape[0]


def enc_bufs(nc: int, nf: int = 64):
    """Channel counts of the encoder's activation buffers per level (A: H, B: H/2, C: H/4)."""
    return {"A": {0: nc, 1: nc, 2: nf}, "B": {0: nf, 1: nf, 2: 2 * nf}, "C": {0: 2 * nf, 1: 2 * nf, 2: 2 * nf, 3: 2 * nf, 4: 4 * nf, 5: nc}}


def dec_bufs(nc: int, nf: int = 64):
    return {"C": {0: nc, 1: 4 * nf, 2: 4 * nf, 3: 2 * nf, 4: 2 * nf, 5: 2 * nf, 6: 2 * nf}, "B": {0: 2 * nf, 1: 2 * nf, 2: nf},
            "A": {0: nf, 1: nf, 2: nf, 3: nf, 4: nc}}


ENC_BUFS, DEC_BUFS = enc_bufs(64), dec_bufs(64)
ENC_IN, ENC_OUT = ("A", 0), ("C", 5)
DEC_IN, DEC_OUT = ("C", 0), ("A", 4)


def encoder_graph(sd, e="srvp_encoder."):
    """SmallEncoder.forward (res_models.py:96-109) as a list of ops over three resolution levels A (H), B (H/2), C (H/4):
    ("stage", level, StageDef) | ("pool", (level, buf), (level, buf), channels).  SmallEncoder(nc, nh = nc, nf)."""
    nc, nf = codec_widths(sd, e)
    g = [("stage", "A", s) for s in _res_block("enc0", sd, e + "blocks.0", 0, 1, 2, nc, nf)]
    g.append(("pool", ("A", 2), ("B", 0), nf))
    g += [("stage", "B", s) for s in _res_block("enc1", sd, e + "blocks.1", 0, 1, 2, nf, 2 * nf)]
    g.append(("pool", ("B", 2), ("C", 0), 2 * nf))
    st = _res_block("enc2", sd, e + "blocks.2", 0, 1, 2, 2 * nf, 2 * nf) + _res_block("enc3", sd, e + "blocks.3", 2, 1, 3, 2 * nf, 2 * nf)
    st += _res_block("enc4", sd, e + "blocks.4", 3, 1, 4, 2 * nf, 4 * nf)
    st += _conv_bn_act("enc_last", sd, e + "last_conv.0", 4, 5, act=L.ACT_TANH)
    return g + [("stage", "C", s) for s in st]


def decoder_graph(sd, d="srvp_decoder."):
    """SmallDecoder.forward (res_models.py:134-147, skip=None): ("stage", ...) | ("up", (level, buf), (level, buf), channels)."""
    nf = sd[d + "blocks.4.layers.conv_2.conv.weight"].shape[0]
    st = _conv_bn_act("dec_first", sd, d + "first_upconv", 0, 1, w=_convT_as_conv(sd[d + "first_upconv.conv.weight"].float()))
    st += _res_block("dec0", sd, d + "blocks.0", 1, 2, 3, 4 * nf, 2 * nf)
    st += _res_block("dec1", sd, d + "blocks.1", 3, 4, 5, 2 * nf, 2 * nf) + _res_block("dec2", sd, d + "blocks.2", 5, 4, 6, 2 * nf, 2 * nf)
    g = [("stage", "C", s) for s in st]
    g.append(("up", ("C", 6), ("B", 0), 2 * nf))
    g += [("stage", "B", s) for s in _res_block("dec3", sd, d + "blocks.3", 0, 1, 2, 2 * nf, nf)]
    g.append(("up", ("B", 2), ("A", 0), nf))
    st = _res_block("dec4", sd, d + "blocks.4", 0, 1, 2, nf, nf) + _conv_bn_act("dec_last0", sd, d + "last_conv.0", 2, 3)
    wl = _convT_as_conv(sd[d + "last_conv.1.conv.weight"].float())
    st.append(StageDef("dec_last1", L.EPI_BIAS_LRELU, sd[d + "last_conv.1.conv.bias"].float(), [4], flags=_act_flags(L.ACT_LRELU, out32=True))
              .add(3, wl, 0, 1))
    return g + [("stage", "A", s) for s in st]


class CodecEngine:
    """SmallEncoder + SmallDecoder of one NNFOwithBayesianJumps for a fixed BEV size: input = latent width nc = nh =
    MODEL.ENCODER.OUT_CHANNELS = 64 or 128 (BASELINE config 5), filter size nf = 64 (the reference's default, config.py:115) or 128.
    SKIPCO = True cannot run in the reference either (SmallDecoder.forward asserts on skip=None, which is what
    temporal_ode_bayes.py:624 passes; with skips its last_conv would see nf instead of 2 nf channels), so there is nothing to mirror."""

    def __init__(self, sd: Dict[str, torch.Tensor], H: int, W: int, n_enc: int, n_dec: int, precision: str, device):
        if H % 4 or W % 4:
            raise L.SfError("BEV height / width must be multiples of 4 for the fused encoder / decoder")
        self.lib = L.load()
        self.H, self.W, self.h, self.w = H, W, H // 4, W // 4
        self.device = device
        # The encoder always runs in the accurate split-bf16 mode: its output feeds every observation jump, and bf16 operands
        # through its 11 convolutions would alone spend the hidden state's 1e-2 budget (measured 1.5e-2 on the reference
        # fixtures).  The decoder only shapes the output frames and follows the requested precision.
        self.enc_x3 = True
        self.x3 = precision == "bf16x3"
        self.n_enc, self.n_dec = n_enc, n_dec
        sd = {k: v.detach().to(device) for k, v in sd.items() if k.startswith(("srvp_encoder", "srvp_decoder"))}
        nc, nf = codec_widths(sd)
        self.nc, self.nf = nc, nf
        if (nc not in (64, 128) or nf not in (64, 128) or sd["srvp_encoder.last_conv.0.conv.weight"].shape[0] != nc
                or sd["srvp_decoder.first_upconv.conv.weight"].shape[0] != nc
                or sd["srvp_decoder.blocks.0.layers.conv_1.conv.weight"].shape[0] != 4 * nf):
            raise L.SfError("fused encoder / decoder is built for nc = nh in (64, 128), nf in (64, 128), SKIPCO off")
        self.launches = 0
        dims = {"A": (H, W), "B": (H // 2, W // 2), "C": (H // 4, W // 4)}
        self.dims = dims
        self.enc, self.enc_ops = self._build(encoder_graph(sd), enc_bufs(nc, nf), n_enc, dims, self.enc_x3)
        self.dec, self.dec_ops = self._build(decoder_graph(sd), dec_bufs(nc, nf), n_dec, dims, self.x3)
        self.dec_out32 = torch.empty((n_dec, H, W, nc), dtype=torch.float32, device=device)
        self.dec["A"].bind_out32(self.dec_out32)

    def _build(self, graph, bufs, n, dims, x3):
        plans = {}
        for lvl, chans in bufs.items():
            plans[lvl] = LevelPlan(self.lib, dims[lvl][0], dims[lvl][1], n, x3, self.device)
            for b, ch in chans.items():
                plans[lvl].buf(b, ch)
        ops = []
        for op in graph:
            if op[0] == "stage":
                slot = plans[op[1]].stage(op[2])
                if ops and ops[-1][0] == "stages" and ops[-1][1] == op[1]:
                    ops[-1][2].append(slot)
                else:
                    ops.append(("stages", op[1], [slot]))
            else:
                ops.append(op)
        for p in plans.values():
            p.finalize()
        return plans, ops

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _run(self, plans, ops, n):
        k = 0
        for op in ops:
            if op[0] == "stages":
                k += plans[op[1]].run(op[2], n)
                continue
            (ls, bs), (ld, bd), ch = op[1], op[2], op[3]
            src, dst = plans[ls].bufs[bs], plans[ld].bufs[bd]
            H, W = self.dims[ls]
            if op[0] == "pool":
                L.check(self.lib.sf_maxpool2(src[0].data_ptr(), src[1].data_ptr() if src[1] is not None else None, dst[0].data_ptr(),
                                             dst[1].data_ptr() if dst[1] is not None else None, n, H, W, ch, self._stream()), "maxpool")
                k += 1
            else:
                for s_, d_ in zip(src, dst):
                    if s_ is not None:
                        L.check(self.lib.sf_upsample2(s_.data_ptr(), d_.data_ptr(), n, H, W, ch, self._stream()), "upsample")
                        k += 1
        return k

    def encode(self, frames_nchw: torch.Tensor):
        """frames [n, nc, H, W] fp32 NCHW -> encoded latents as NHWC bf16 planes (hi, lo) [n, h, w, nc] (views of the engine's
        output buffer: valid until the next encode)."""
        n = frames_nchw.shape[0]
        assert n <= self.n_enc and tuple(frames_nchw.shape[1:]) == (self.nc, self.H, self.W)
        src = frames_nchw.contiguous().float()
        a_in = self.enc[ENC_IN[0]].bufs[ENC_IN[1]]
        L.check(self.lib.sf_pack_nchw_f32(src.data_ptr(), a_in[0].data_ptr(), a_in[1].data_ptr() if a_in[1] is not None else None, n, self.nc,
                                          self.H, self.W, self._stream()), "pack")
        self.launches += 1 + self._run(self.enc, self.enc_ops, n)
        hi, lo = self.enc[ENC_OUT[0]].bufs[ENC_OUT[1]]
        return hi[:n], (lo[:n] if lo is not None else None)

    def decode(self, path_nhwc_f32: torch.Tensor, slots_dev: torch.Tensor, unpack: bool = True):
        """Recorded latent states (the ODE engine's fp32 NHWC path buffer), gathered by slot -> decoded frames
        [n, nc, H, W] fp32 NCHW; with unpack=False the engine-layout result instead: ((hi, lo) NHWC bf16 planes, fp32 NHWC),
        views of the decoder's output buffers for the fused refinement."""
        n = slots_dev.numel()
        assert n <= self.n_dec
        z = self.dec[DEC_IN[0]].bufs[DEC_IN[1]]
        L.check(self.lib.sf_cast_nhwc_f32(path_nhwc_f32.data_ptr(), slots_dev.data_ptr(), z[0].data_ptr(), z[1].data_ptr() if z[1] is not None else None,
                                          n, self.nc, self.h, self.w, self._stream()), "cast")
        k = self._run(self.dec, self.dec_ops, n)
        if not unpack:
            self.launches += k + 1
            hi, lo = self.dec[DEC_OUT[0]].bufs[DEC_OUT[1]]
            return (hi[:n], lo[:n] if lo is not None else None), self.dec_out32[:n]
        out = torch.empty((n, self.nc, self.H, self.W), dtype=torch.float32, device=self.device)
        L.check(self.lib.sf_unpack_nhwc_f32(self.dec_out32.data_ptr(), out.data_ptr(), None, n, self.nc, self.H, self.W, self._stream()), "unpack")
        self.launches += k + 2
        return out
