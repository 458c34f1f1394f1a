"""GPU parity of the BEV Decoder head (streamingflow/models/decoder.py:91-140 + the arg-max of
trainer.py:230-231) on the CUDA engine: reference-run fixtures, the fp64 oracle at the BEV size, and the whole
ODE head -> Decoder -> masks pipeline against the masks the unmodified reference produced."""
import os

import numpy as np
import pytest
import torch

from oracle import sf_oracle as so
from oracle._refimport import make_cfg

pytestmark = pytest.mark.gpu
TOL = {"bf16": 1e-2, "bf16x3": 1e-4}
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ALL_GATES = dict(perceive_hdmap=True, predict_pedestrian=True, predict_instance=True, predict_future_flow=True, planning=True)
SEG_ONLY = {k: False for k in ALL_GATES}


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12)).item()


def _decoder(gates, n_present, seed, gain, precision):
    from streamingflow_b200.models.decoder import Decoder

    m = Decoder(64, 2, n_present, 2, gates).eval()
    m.load_state_dict(so.recipe_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed, gain), strict=True)
    m.precision = precision
    return m.cuda()


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_decoder_heads_match_reference_fixture(precision):
    """Every head of the Decoder (all predict gates on) vs the outputs of the UNMODIFIED reference (decoder_all_c64.npz)."""
    z = np.load(os.path.join(GOLDEN, "decoder_all_c64.npz"))
    m = _decoder(ALL_GATES, int(z["n_present"]), int(z["seed"]), float(z["gain"]), precision)
    x = so.recipe_array("dec_in", (1, 3, 64, 32, 32), int(z["seed"])).cuda()
    with torch.no_grad():
        out = m(x)
    torch.cuda.synchronize()
    for k in ("segmentation", "pedestrian", "hdmap", "instance_center", "instance_offset", "instance_flow", "costvolume"):
        want = torch.from_numpy(z[k])
        assert out[k].shape == want.shape, (k, out[k].shape, want.shape)
        assert _rel(out[k].cpu(), want) < 2 * TOL[precision], f"{k}: {_rel(out[k].cpu(), want):.3e} ({precision})"
    assert torch.equal(out["segmentation_argmax"].long(), out["segmentation"].argmax(dim=2))      # the fused arg-max == torch's
    if precision == "bf16x3":
        assert torch.equal(out["segmentation"].argmax(dim=2).cpu(), torch.from_numpy(z["segmentation"]).argmax(dim=2))


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_decoder_matches_oracle_at_bev_200(precision):
    """Segmentation branch at the real BEV size (200 x 200, 4 frames: levels 200 / 100 / 50 / 25, ragged 16 x 8 tiles at 25 x 25)
    vs the fp64 oracle on the GPU; NCHW fp32 input and engine-layout planes give the same result."""
    m = _decoder(SEG_ONLY, 3, 17, 1.0, precision)
    x = so.recipe_array("dec_big", (2, 2, 64, 200, 200), 17).cuda()
    with torch.no_grad():
        out = m(x)
    sd = {"d." + k: (v.double() if v.is_floating_point() else v) for k, v in m.state_dict().items()}
    with torch.no_grad():
        want = so.seg_decoder(sd, "d", x.double())
    assert _rel(out["segmentation"], want) < TOL[precision], _rel(out["segmentation"], want)
    flips = out["segmentation_argmax"].long() != want.argmax(2)
    margin = (want[:, :, 0] - want[:, :, 1]).abs()
    dmax = (out["segmentation"].double() - want).abs().max()
    assert not bool((flips & (margin > 2 * dmax)).any())
    # engine-layout input: the same frames as NHWC bf16 planes (what the fused refinement hands over)
    nhwc = x.view(4, 64, 200, 200).permute(0, 2, 3, 1).contiguous()
    hi = nhwc.to(torch.bfloat16)
    lo = (nhwc - hi.float()).to(torch.bfloat16)
    with torch.no_grad():
        out2 = m(x, planes=(hi, lo if precision == "bf16x3" else None))
    assert torch.equal(out2["segmentation"], out["segmentation"]) and torch.equal(out2["segmentation_argmax"], out["segmentation_argmax"])


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_ode_head_to_masks_pipeline_matches_reference_masks(precision):
    """FuturePredictionODE.forward -> Decoder -> arg-max, every stage on the CUDA engine and the frames handed over in engine
    layout (no [B, T, 64, 200, 200] fp32 round trip), vs the masks the unmodified reference produced (argmax_c64.npz, the seeds
    with the widest smallest logit margin).  Accurate mode: bit-exact.  bf16: only pixels whose reference margin is below
    twice the measured logit error may differ."""
    from streamingflow_b200.models.decoder import Decoder
    from streamingflow_b200.models.future_prediction_ode import FuturePredictionODE

    z = np.load(os.path.join(GOLDEN, "argmax_c64.npz"))
    C, H, gain = 64, int(z["H"]), float(z["gain"])
    ct = torch.tensor([[-1.0, -0.5, 0.0]], dtype=torch.float64)
    tt = torch.tensor([[0.5, 1.0, 1.5, 2.0]], dtype=torch.float64)
    for k, seed in enumerate(int(s) for s in z["seeds"]):
        want = torch.from_numpy(np.unpackbits(z["masks"][k], axis=-1)[..., :H].astype(np.int64))
        m = FuturePredictionODE(C, C, 4, make_cfg(C)).eval()
        sd32 = so.recipe_state_dict({kk: tuple(v.shape) for kk, v in m.state_dict().items()}, seed, gain)
        m.load_state_dict(sd32, strict=True)
        m = m.cuda()
        m.gru_ode.precision = precision
        d = Decoder(C, 2, 3, 2, SEG_ONLY).eval()
        dsd32 = so.recipe_state_dict({kk: tuple(v.shape) for kk, v in d.state_dict().items()}, seed, gain)
        d.load_state_dict(dsd32, strict=True)
        d = d.cuda()
        d.precision = precision
        cam = so.recipe_array("cam", (1, 3, C, H, H), seed).cuda()
        tape = torch.stack([so.recipe_array(f"eps{i}", (C, H // 4, H // 4), seed) for i in range(int(z["n_eps"][k]))]).cuda()
        m.gru_ode._draw_noise = lambda n, h, w, device, _t=tape: _t[:n].contiguous()
        with torch.no_grad():
            x, _ = m(torch.zeros(1, 1, C, H, H, device="cuda"), cam, None, ct, None, tt)
            assert m.last_output_planes is not None
            out = d(x, planes=m.last_output_planes)
            # fp64 oracle logits for the margins
            sd64 = {kk: (v.double().cuda() if v.is_floating_point() else v.cuda()) for kk, v in sd32.items()}
            dsd64 = {"d." + kk: (v.double().cuda() if v.is_floating_point() else v.cuda()) for kk, v in dsd32.items()}
            xo = so.future_prediction_forward(sd64, cam.double(), None, ct, None, tt, 0.05, iter(tape.double()[:, None]))
            seg_ref = so.seg_decoder(dsd64, "d", xo)
        got = out["segmentation_argmax"].long().cpu()
        assert torch.equal(seg_ref.argmax(2).cpu(), want)
        flips = got != want
        margin = (seg_ref[:, :, 0] - seg_ref[:, :, 1]).abs().cpu()
        dmax = float((out["segmentation"].double() - seg_ref).abs().max())
        info = dict(seed=seed, precision=precision, flips=int(flips.sum()), logit_err=dmax, logit_rel_err=dmax / float(seg_ref.abs().max()),
                    min_margin=float(margin.min()))
        assert info["logit_rel_err"] < 5 * TOL[precision], info
        if precision == "bf16x3":
            assert torch.equal(got, want), info
        else:
            assert not bool((flips & (margin > 2 * dmax)).any()), info
