"""Pins the CPU oracle (oracle/sf_oracle.py) against fixtures produced by RUNNING THE UNMODIFIED REFERENCE
(oracle/gen_golden.py; reference files cited per function in the oracle).  CPU-only, runs in seconds."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import sf_oracle as so


def _shapes(z):
    return {k: tuple(int(s) for s in v.split(",") if s) for k, v in zip(z["shapes_keys"], z["shapes_vals"])}


def test_schedule_traces_match_reference(golden_dir):
    """temporal_ode_bayes.py:508,539-553,585-604 (schedule) and :606-622 (selection), incl. 1-ulp micro-steps."""
    cases = json.load(open(os.path.join(golden_dir, "sched.json")))["cases"]
    assert len(cases) >= 15
    n_micro = 0
    for c in cases:
        sch = so.build_schedule(c["times"], c["targets"], c["delta_t"], c["variable"])
        evs = sch.events
        if c["solver"] == "midpoint":       # the reference trace counts ode_step calls, same as ours
            pass
        assert [e.kind for e in evs] == c["kinds"], c["tag"]
        for e, dt, ta in zip(evs, c["dts"], c["t_after"]):
            if e.kind == "step":
                assert e.dt == dt, (c["tag"], e.dt, dt)               # bit-exact doubles
                assert e.t_after == ta, (c["tag"], e.t_after, ta)
        assert [sch.path_ev[i] for i in sch.select] == c["selected"], c["tag"]
        n_micro += any(e.kind == "step" and e.dt < 1e-9 for e in evs)
    assert n_micro >= 3, "fixture must contain micro-step cases (SURVEY F6)"


def test_schedule_with_float32_stamps_matches_reference_traces(golden_dir):
    """Reference traces with float32 timestamp tensors (tests/golden/sched_f32.json): comparisons and variable steps are
    float32 arithmetic there; most of these schedules differ from the float64 arithmetic on the same stamps."""
    cases = json.load(open(os.path.join(golden_dir, "sched_f32.json")))["cases"]
    differ = 0
    for c in cases:
        sch = so.build_schedule(c["times"], c["targets"], c["delta_t"], c["variable"], stamp_dtype=np.float32)
        assert [e.kind for e in sch.events] == c["kinds"], c["tag"]
        for e, dt, ta in zip(sch.events, c["dts"], c["t_after"]):
            if e.kind == "step":
                assert e.dt == dt and e.t_after == ta, (c["tag"], e.dt, dt, e.t_after, ta)
        assert [sch.path_ev[i] for i in sch.select] == c["selected"], c["tag"]
        s64 = so.build_schedule(c["times"], c["targets"], c["delta_t"], c["variable"])
        differ += [(e.kind, e.dt) for e in s64.events] != [(e.kind, e.dt) for e in sch.events]
    assert differ >= 10


def test_known_answer_counts():
    """SURVEY 8(c) known-answer schedule counts, probed on the reference."""
    times = sorted([-1.0, -0.5, 0.0, -0.8, -0.6, -0.4, -0.2, 0.0])
    tg = [-1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0]
    v = so.build_schedule(times, tg, 0.05, True)
    assert sum(e.kind == "jump" for e in v.events) == 8 and sum(e.kind == "step" for e in v.events) == 10
    f = so.build_schedule(times, tg, 0.05, False)
    assert sum(e.kind == "step" for e in f.events) == 60


def test_sort_is_stable_camera_first():
    order = so.sort_observations([-1.0, -0.5, 0.0], [-0.8, -0.5, 0.0])
    assert [(s, i) for _, s, i in order] == [("cam", 0), ("lidar", 0), ("cam", 1), ("lidar", 1), ("cam", 2), ("lidar", 2)]


def test_full_module_matches_reference(golden_dir):
    """FuturePredictionODE.forward end to end (encoder, jump/ODE loop, selection, decoder, SpatialGRU x2,
    ConvNeXt block, DeepLabHead) at C=8, B=2 with jittered stamps; fp64 -> exact to rounding."""
    z = np.load(os.path.join(golden_dir, "tiny_full_c8.npz"))
    C, H, B, seed = int(z["C"]), int(z["H"]), int(z["B"]), int(z["seed"])
    for dtype, tol in ((torch.float64, 1e-12), (torch.float32, 2e-5)):
        sd = so.recipe_state_dict(_shapes(z), seed, float(z["gain"]), dtype)
        cam = so.recipe_array("cam", (B, 3, C, H, H), seed, dtype)
        lid = so.recipe_array("lidar", (B, 5, C, H, H), seed, dtype)
        eps = iter([so.recipe_array(f"eps{i}", (1, C, H // 4, H // 4), seed, dtype) for i in range(64)])
        with torch.no_grad():
            x = so.future_prediction_forward(sd, cam, lid, torch.from_numpy(z["camera_timestamp"]),
                                             torch.from_numpy(z["lidar_timestamp"]), torch.from_numpy(z["target_timestamp"]),
                                             0.05, eps)
        ref = torch.from_numpy(z["x_f64"])
        err = ((x.double() - ref).abs().max() / ref.abs().max()).item()
        assert err < tol, (dtype, err)
        assert sum(1 for _ in eps) == 64 - int(z["n_eps"])      # same number of noise draws as the reference


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "c64_latent_*.npz")) +
                                        glob.glob(os.path.join(os.path.dirname(__file__), "golden", "c128_latent_*.npz"))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_c64_latent_rollout_matches_reference(path):
    """NNFOwithBayesianJumps.forward at the production width C=64 (and the 128-channel network of config 5): per-event
    latent states, selection, decode."""
    z = np.load(path)
    C, H, seed = int(z["C"]), int(z["H"]), int(z["seed"])
    from oracle.shapes import nnfo_shapes

    dtype = torch.float64
    sd = {"g." + k: v for k, v in so.recipe_state_dict(nnfo_shapes(C), seed, float(z["gain"]), dtype).items()}
    times, targets = z["times"].tolist(), z["targets"].tolist()
    obs = so.recipe_array("obs", (1, len(times), C, H, H), seed, dtype)
    eps = iter([so.recipe_array(f"eps{i}", (1, C, H // 4, H // 4), seed, dtype) for i in range(256)])
    trace = []
    with torch.no_grad():
        state, sel, x = so.nnfo_forward(sd, "g", times, obs, targets, 0.05, eps, solver=str(z["solver"]),
                                        impute=bool(z["impute"]), variable=bool(z["variable"]), trace=trace)
    ref_states = torch.from_numpy(z["states_f64"]).double()
    got = torch.cat(trace, 0)
    assert got.shape == ref_states.shape
    assert ((got - ref_states).abs().max() / ref_states.abs().max()).item() < 2e-7      # fixture stored as fp32
    assert ((state - torch.from_numpy(z["final_f64"]).double()).abs().max()).item() < 1e-6
    xr = torch.from_numpy(z["x_f64"]).double()
    assert ((x[:, [0, -1]] - xr).abs().max() / xr.abs().max()).item() < 2e-7
    sch = so.build_schedule(times, targets, 0.05, bool(z["variable"]))
    assert [sch.path_ev[i] for i in sch.select] == z["selected"].tolist()
    assert sum(1 for _ in eps) == 256 - int(z["n_eps"])


def test_operand_rounding_budget():
    """Error budget for the CUDA precision modes, emulated on the oracle (no GPU): plain tf32 operands cannot
    reach 1e-4 on this rollout, the split-bf16 (3-product) path can; bf16 stays under 1e-2."""
    from oracle.shapes import nnfo_shapes

    C, H, seed = 64, 16, 7
    times = sorted([-1.0, -0.5, 0.0, -0.8, -0.6, -0.4, -0.2, 0.0])
    sch = so.build_schedule(times, [-1.0, 0.0, 1.0, 2.0], 0.05, True)

    def run(dtype, ctx):
        sd = {"g." + k: v for k, v in so.recipe_state_dict(nnfo_shapes(C), seed, 1.0, dtype).items()}
        hx = torch.tanh(so.recipe_array("hx", (len(times), C, H // 4, H // 4), seed, dtype))
        eps = iter([so.recipe_array(f"eps{i}", (1, C, H // 4, H // 4), seed, dtype) for i in range(64)])
        tr = []
        with torch.no_grad():
            if ctx is None:
                so.integrate_latent(sd, "g", hx, sch, eps, trace=tr)
            else:
                with ctx:
                    so.integrate_latent(sd, "g", hx, sch, eps, trace=tr)
        return torch.cat(tr, 0).double()

    truth = run(torch.float64, None)
    err = lambda t: ((t - truth).abs().max() / truth.abs().max()).item()
    e_bf16 = err(run(torch.float32, so.operand_rounding(so.round_bf16)))
    e_tf32 = err(run(torch.float32, so.operand_rounding(so.round_tf32)))
    e_x3 = err(run(torch.float32, so.operand_rounding(so.round_bf16, split3=True)))
    assert e_bf16 < 1e-2 and e_x3 < 1e-4 and e_tf32 > 1e-4, (e_bf16, e_tf32, e_x3)


def test_decoder_segmentation_branch_matches_reference(golden_dir):
    """models/decoder.py (segmentation output) -- used by the GPU tests to turn the ODE head's output into occupancy logits."""
    z = np.load(os.path.join(golden_dir, "decoder_seg_c64.npz"))
    sd = {"d." + k: v for k, v in so.recipe_state_dict(_shapes(z), int(z["seed"]), float(z["gain"]), torch.float64).items()}
    x = so.recipe_array("dec_in", (1, 2, 64, 32, 32), int(z["seed"]), torch.float64)
    with torch.no_grad():
        seg = so.seg_decoder(sd, "d", x)
    assert (seg - torch.from_numpy(z["seg_f64"])).abs().max().item() < 1e-11


def test_decoder_all_heads_match_reference(golden_dir):
    """models/decoder.py with every predict gate on (reference run: tests/golden/decoder_all_c64.npz): pins oracle.bev_decoder,
    the checker of the CUDA Decoder head."""
    z = np.load(os.path.join(golden_dir, "decoder_all_c64.npz"))
    sd = {"d." + k: v for k, v in so.recipe_state_dict(_shapes(z), int(z["seed"]), float(z["gain"]), torch.float64).items()}
    x = so.recipe_array("dec_in", (1, 3, 64, 32, 32), int(z["seed"]), torch.float64)
    with torch.no_grad():
        out = so.bev_decoder(sd, "d", x, int(z["n_present"]))
    for k in ("segmentation", "pedestrian", "hdmap", "instance_center", "instance_offset", "instance_flow", "costvolume"):
        want = torch.from_numpy(z[k])
        assert out[k].shape == want.shape and (out[k] - want).abs().max().item() < 1e-11, k
