"""CPU tests of the product's host logic: schedule, rollout compilation, weight packing plan, module glue, C-ABI symbols.
No GPU and no CUDA compute: the engine is replaced by a checker backend built on the oracle (tests/_oracle_backend.py)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from oracle import sf_oracle as so
from oracle.shapes import nnfo_shapes
from oracle._refimport import make_cfg
from streamingflow_b200 import schedule as sc
from streamingflow_b200.rollout import compile_rollout
from tests._oracle_backend import OracleBackend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_schedule_matches_reference_traces(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "sched.json")))["cases"]
    for c in cases:
        plan = sc.plan_sample(c["times"], c["targets"], c["delta_t"], c["variable"], c["solver"])
        assert [("step" if o.kind == sc.STEP else "jump") for o in plan.ops] == c["kinds"], c["tag"]
        for o, dt, ta in zip(plan.ops, c["dts"], c["t_after"]):
            if o.kind == sc.STEP:
                assert o.dt == dt and o.t == ta, c["tag"]
        assert plan.picks == c["selected"], c["tag"]


def test_product_schedule_with_float32_stamps_matches_reference_traces(golden_dir):
    """float32 timestamp tensors make the reference's loop compare / subtract in float32; plan_sample reproduces it."""
    cases = json.load(open(os.path.join(golden_dir, "sched_f32.json")))["cases"]
    for c in cases:
        plan = sc.plan_sample(c["times"], c["targets"], c["delta_t"], c["variable"], c["solver"], obs_dtype="float32", target_dtype="float32")
        assert [("step" if o.kind == sc.STEP else "jump") for o in plan.ops] == c["kinds"], c["tag"]
        for o, dt, ta in zip(plan.ops, c["dts"], c["t_after"]):
            if o.kind == sc.STEP:
                assert o.dt == dt and o.t == ta, (c["tag"], o.dt, dt)
        assert plan.picks == c["selected"], c["tag"]


def test_schedule_matches_oracle_on_random_stamps():
    rng = np.random.RandomState(0)
    for trial in range(200):
        n = rng.randint(1, 9)
        times = sorted((rng.uniform(-1.2, 0.05, n)).tolist())
        tg = sorted(rng.uniform(-1.2, 2.5, rng.randint(1, 9)).tolist())
        var = bool(trial & 1)
        a = sc.plan_sample(times, tg, 0.05, var)
        b = so.build_schedule(times, tg, 0.05, var)
        assert [(o.kind == sc.STEP, o.dt) for o in a.ops] == [(e.kind == "step", e.dt) for e in b.events]
        assert a.picks == [b.path_ev[i] for i in b.select]


def test_merge_observations_stable_and_keeps_duplicates():
    order = sc.merge_observations([-1.0, -0.5, 0.0], [-0.8, -0.5, 0.0, 0.0])
    assert [(s, i) for _, s, i in order] == [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2), (1, 2), (1, 3)]
    with pytest.raises(ValueError):
        sc.plan_sample([], [0.0], 0.05, True)


def test_rollout_grouping_and_noise_order():
    times = sorted([-1.0, -0.5, 0.0, -0.8, -0.6, -0.4, -0.2, 0.0])
    tg = [-1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0]
    p0 = sc.plan_sample(times, tg, 0.05, True)
    p1 = sc.plan_sample([t - 0.013 if i % 3 else t for i, t in enumerate(times)], tg, 0.05, True)
    ro = compile_rollout([p0, p1], [0, 8], "euler", True)
    assert ro.n_eps == p0.n_noise + p1.n_noise and ro.n_state_steps == p0.n_steps + p1.n_steps
    seen = {}
    for e in ro.events:
        for b, s in zip(e["samples"], e["eps"]):
            seen.setdefault(b, []).append(s)
    assert seen[0] == list(range(p0.n_noise))                                   # sample-major noise slots
    assert seen[1] == list(range(p0.n_noise, p0.n_noise + p1.n_noise))
    assert all(len(set(e["samples"])) == len(e["samples"]) for e in ro.events)
    rm = compile_rollout([p0], [0], "midpoint", False)
    steps = [e for e in rm.events if e["kind"] == sc.STEP]
    assert len(steps) == 2 * p0.n_steps and rm.n_eps == sc.plan_sample(times, tg, 0.05, True, "midpoint").n_noise
    assert steps[0]["x_buf"] == 4 and steps[0]["run_prior"] == 1 and steps[1]["x_buf"] == 2 and steps[1]["run_prior"] == 0


def test_prior_net_runs_only_where_its_sample_is_read(golden_dir):
    """compile_rollout evaluates the prior net (infer_state) only after ops whose sampled input a following ode_step reads: not
    before a jump (GRUObservationCell ignores p, temporal_ode_bayes.py:327-344), not after the last op; noise slots keep the
    reference's numbering; a streaming push keeps its last input alive; and the module's output (reference fixture) does not
    depend on the switch."""
    times = sorted([-1.0, -0.5, 0.0, -0.8, -0.6, -0.4, -0.2, 0.0])
    tg = [-1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0]
    p = sc.plan_sample(times, tg, 0.05, True)
    kinds = [o.kind for o in p.ops]
    live = [i + 1 < len(kinds) and kinds[i + 1] == sc.STEP for i in range(len(kinds))]
    ro = compile_rollout([p], [0], "euler", True)
    full = compile_rollout([p], [0], "euler", True, skip_dead_prior=False)
    assert [e["run_prior"] for e in ro.events] == [int(v) for v in live] and all(e["run_prior"] == 1 for e in full.events)
    assert ro.n_prior_evals == sum(live) < full.n_prior_evals == len(kinds) == 18
    assert [e["eps"] for e in ro.events] == [e["eps"] for e in full.events] and ro.n_eps == full.n_eps == 18
    assert compile_rollout([p], [0], "euler", True, keep_last_input=True).events[-1]["run_prior"] == 1
    mid = compile_rollout([p], [0], "midpoint", True)
    firsts = [e for e in mid.events if e["kind"] == sc.STEP and e["s_out"] == 1]
    assert firsts and all(e["run_prior"] == 1 for e in firsts)                    # pk = infer_state(k) feeds the second half step
    assert compile_rollout([p], [0], "euler", False).n_prior_evals == 0
    # the module on the reference's fixture, both settings: identical output
    z = np.load(os.path.join(golden_dir, "tiny_full_c8.npz"))
    C, H, B, seed = int(z["C"]), int(z["H"]), int(z["B"]), int(z["seed"])
    outs = []
    for skip in (True, False):
        m = _tiny_module(z, torch.float64)
        m.gru_ode.skip_dead_prior = skip
        m.gru_ode._draw_noise = lambda n, h, w, device: torch.stack(
            [so.recipe_array(f"eps{i}", (C, H // 4, H // 4), seed, torch.float64) for i in range(n)])
        with torch.no_grad():
            x, _ = m(torch.zeros(B, 1, C, H, H, dtype=torch.float64), so.recipe_array("cam", (B, 3, C, H, H), seed, torch.float64),
                     so.recipe_array("lidar", (B, 5, C, H, H), seed, torch.float64), torch.from_numpy(z["camera_timestamp"]),
                     torch.from_numpy(z["lidar_timestamp"]), torch.from_numpy(z["target_timestamp"]))
        outs.append((x, m.gru_ode.last_rollout.n_prior_evals))
    assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1] < outs[1][1]
    ref = torch.from_numpy(z["x_f64"])
    assert ((outs[0][0] - ref).abs().max() / ref.abs().max()).item() < 1e-11


@pytest.mark.parametrize("x3", [False, True])
def test_packed_stage_plans_reproduce_the_convolutions(x3):
    """The chunk / tap / column plan + packed weight matrix the TMA ring streams, replayed on the host, equals F.conv2d."""
    from streamingflow_b200 import engine as en
    import torch.nn.functional as F

    torch.manual_seed(0)
    sd = {k: v for k, v in so.recipe_state_dict(nnfo_shapes(64), 5, 1.0, torch.float32).items()}
    H, W = 5, 6
    rnd = lambda c: torch.randn(1, c, H, W)
    q = (lambda t: t) if x3 else (lambda t: t.to(torch.bfloat16).float())
    tol = 2e-4 if x3 else 1e-12
    x, s, g1, g2, a, b = (rnd(64) for _ in range(6))
    src = {-1: x, -2: s, -3: s, en.BUF_G1: g1, en.BUF_G2: g2, en.BUF_A: a, en.BUF_B: b, en.BUF_HH: a, en.BUF_T1: g1,
           en.BUF_T2: g2, en.BUF_Q1: x, en.BUF_Y1: rnd(128), en.BUF_Q3: rnd(128), en.BUF_Y2: rnd(128)}
    cell = en.cell_stage_defs(sd, "gru_c")
    conv = lambda inp, key, pad: F.conv2d(q(inp).double(), q(sd[key]).double(), None, padding=pad)[0]
    if x3:   # compare against exact fp64 convs; the split carries ~2^-16 relative error
        conv = lambda inp, key, pad: F.conv2d(inp.double(), sd[key].double(), None, padding=pad)[0]
    xs, ss = torch.cat([x, s], 1), torch.cat([s, s], 1)
    acc = en.emulate_stage(cell[0], x3, src)
    ref = torch.cat([conv(xs, "gru_c.conv_update_1.weight", 1), conv(xs, "gru_c.conv_reset_1.weight", 1),
                     conv(ss, "gru_c.conv_update_2.weight", 1), conv(ss, "gru_c.conv_reset_2.weight", 1)])
    # the cat[s,s] fold rounds (Wa+Wb) once instead of Wa and Wb separately: compare u2/r2 loosely in bf16 mode
    assert (acc[:128] - ref[:128]).abs().max() < max(tol, 1e-9) * 10 + (0 if x3 else 0)
    assert (acc[128:] - ref[128:]).abs().max() < (2e-4 if x3 else 5e-2)
    acc = en.emulate_stage(cell[1], x3, src)
    ref = torch.cat([conv(torch.cat([x, g1], 1), "gru_c.conv_state_tilde_1.weight", 1),
                     conv(torch.cat([s, g2], 1), "gru_c.conv_state_tilde_2.weight", 1)])
    assert (acc[:128] - ref).abs().max() < max(tol, 1e-9) * 10
    by_name = {d.name: d for d in cell}
    assert list(by_name) == list(en.CELL_STAGE_NAMES[64]) == ["gates", "propose", "decode", "trunk", "mix"]
    acc = en.emulate_stage(by_name["trunk"], x3, src)
    ref = conv(torch.cat([a, b], 1), "gru_c.trusting_gate.0.layers.0.weight", 3)
    assert (acc[:64] - ref).abs().max() < max(tol, 1e-9) * 30
    # the fused 1x1 follow-up conv: its [n, k] weights are the last 64 (x3: 128 = hi, lo) rows of the packed matrix, and the
    # stage vector is [LN1 w, LN1 b, LN2 w, LN2 b]; the unfused pair of stages (SF_B2B=0) packs the same 7x7 rows
    _, wp = en.pack_stage(by_name["trunk"], x3)
    w1 = sd["gru_c.trusting_gate.0.layers.3.weight"][:, :, 0, 0]
    tail = wp[-(128 if x3 else 64):].float()
    assert torch.equal(tail[:64], w1.to(torch.bfloat16).float())
    if x3:
        assert torch.equal(tail[64:], (w1 - w1.to(torch.bfloat16).float()).to(torch.bfloat16).float())
    split = {d.name: d for d in en.cell_stage_defs(sd, "gru_c", b2b=False)}
    assert "trunk7" in split and "trunk1" in split and torch.equal(en.pack_stage(split["trunk7"], x3)[1], wp[:-(128 if x3 else 64)])
    assert torch.equal(by_name["trunk"].vec, torch.cat([split["trunk7"].vec, split["trunk1"].vec])) and by_name["trunk"].io == split["trunk1"].io
    acc = en.emulate_stage(by_name["mix"], x3, src)
    ref = torch.cat([conv(g2, "gru_c.trusting_gate.0.layers.6.weight", 1),
                     conv(torch.cat([a, b], 1), "gru_c.trusting_gate.0.projection.0.weight", 0)])
    assert (acc[:128] - ref).abs().max() < max(tol, 1e-9) * 10
    prior = en.prior_stage_defs(sd, "p_model")
    acc = en.emulate_stage(prior[4], x3, src)           # q4: 128 -> 128 with BN folded
    assert prior[2] == 'se0' and prior[5] == 'se1' and prior[4].name == 'q4'
    w4, b4 = en._bn_fold(sd, "p_model.model.2.layers.conv_2")
    ref = F.conv2d((src[en.BUF_Q3] if x3 else q(src[en.BUF_Q3])).double(), (w4 if x3 else q(w4)).double(), None, padding=1)[0]
    assert (acc[:128] - ref).abs().max() < max(tol, 1e-9) * 30
    bn = so._bn_eval(sd, "p_model.model.2.layers.conv_2.norm", F.conv2d(src[en.BUF_Q3], sd["p_model.model.2.layers.conv_2.conv.weight"], None, padding=1))
    assert (F.conv2d(src[en.BUF_Q3], w4, b4, padding=1) - bn).abs().max() < 1e-4


@pytest.mark.parametrize("x3", [False, True])
def test_se_fold_master_matches_packing_of_scaled_weights(x3):
    """SE layer folded into a consumer: se_fold_kernel's arithmetic (fp32 master row x scale[c0 + k] -> bf16, residual rows get the
    rounding residual) replayed on the host equals pack_stage of the conv whose input channels were scaled."""
    from streamingflow_b200 import engine as en

    torch.manual_seed(3)
    sd = so.recipe_state_dict(nnfo_shapes(64), 7, 1.0, torch.float32)
    prior = en.prior_stage_defs(sd, "p_model", fold_se=True)
    q3, q4, q5 = prior[3], prior[4], prior[6]
    assert q3.name == "q3" and q3.fold_se == 0 and q5.fold_se == 1 and q4.fold_se is None
    assert q3.chunks[0][0] == en.BUF_Z1 and q5.chunks[0][0] == en.BUF_Z2 and q4.io[0] == en.BUF_Z1 and q4.flags & 128
    scale = torch.rand(128) + 0.25
    for sdef in (q3, q5):
        w32, meta = en.pack_stage_master(sdef, x3)
        chunks, wp = en.pack_stage(sdef, x3)
        assert w32.shape == wp.shape and meta.shape[0] == wp.shape[0]
        c0 = (meta & 0xffff).long()
        v = w32 * scale[c0[:, None] + torch.arange(64)[None, :]]
        hi = v.to(torch.bfloat16)
        folded = torch.where(((meta >> 16) & 1).bool()[:, None], (v - hi.float()).to(torch.bfloat16), hi)
        scaled = en.StageDef(sdef.name, sdef.epilogue, sdef.vec, sdef.io)
        for buf, cc, w, col, init, ox, oy in sdef.chunks:
            scaled.chunks.append((buf, cc, w * scale[cc:cc + 64][None, :, None, None], col, init, ox, oy))
        _, want = en.pack_stage(scaled, x3)
        assert torch.equal(folded.view(torch.int16), want.view(torch.int16))
    # without the fold the stages read the materialised SE outputs
    plain = en.prior_stage_defs(sd, "p_model")
    assert plain[3].chunks[0][0] == en.BUF_Y1 and plain[6].chunks[0][0] == en.BUF_Y2 and plain[3].fold_se is None


def test_row_paired_tap_packing_order_and_event_group_cap():
    """Host logic of two kernel-side schemes: the dy order of row-paired taps (upper tap of each pair first, odd tap last), the
    rows it produces, and the optional cap on the samples of a batched event (same ops, same noise slots, smaller groups)."""
    from streamingflow_b200 import engine as en
    from streamingflow_b200 import _lib as L

    assert en._pair_order(7) == [1, 0, 3, 2, 5, 4, 6] and en._pair_order(3) == [1, 0, 2] and en._pair_order(2) == [1, 0]
    w = torch.arange(64 * 64 * 7 * 7, dtype=torch.float32).reshape(64, 64, 7, 7) / 4096.0
    sdef = en.StageDef("t", L.EPI_LNGELU, torch.zeros(128), [en.BUF_T1], flags=L.FLAG_PAIR_ROWS).add(en.BUF_A, w, 0, 1)
    chunks, wp = en.pack_stage(sdef, False)
    assert len(chunks) == 1 and chunks[0]["nrep"] == 1 and wp.shape == (49 * 64, 64)
    for dx in (0, 3, 6):
        for slot, dy in enumerate(en._pair_order(7)):
            rows = wp[(dx * 7 + slot) * 64:(dx * 7 + slot + 1) * 64].float()
            assert torch.equal(rows, w[:, :, dy, dx].to(torch.bfloat16).float())
    chunks3, wp3 = en.pack_stage(sdef, True)          # split mode: [pair][rep (hi, lo)][tap] rows, then the lo-plane chunk
    assert [c["nrep"] for c in chunks3] == [2, 1] and wp3.shape[0] == 49 * 64 * 3
    hi = w[:, :, 1, 0].to(torch.bfloat16)
    assert torch.equal(wp3[:64], hi) and torch.equal(wp3[64:128], w[:, :, 0, 0].to(torch.bfloat16))                 # rep hi: dy 1 | dy 0
    assert torch.equal(wp3[128:192], (w[:, :, 1, 0] - hi.float()).to(torch.bfloat16))                                 # rep lo: dy 1 ...

    times, tg = [-1.0, -0.5, 0.0], [-1.0, -0.5, 0.0, 0.5, 1.0]
    plans = [sc.plan_sample(times, tg, 0.05, True, "euler") for _ in range(5)]
    whole = compile_rollout(plans, [0, 3, 6, 9, 12], "euler", True)
    capped = compile_rollout(plans, [0, 3, 6, 9, 12], "euler", True, max_group=2)
    assert all(len(e["samples"]) <= 2 for e in capped.events) and len(capped.events) == 3 * len(whole.events)
    assert capped.n_eps == whole.n_eps and capped.n_state_steps == whole.n_state_steps and capped.out_slots == whole.out_slots
    flat = lambda ro, k: sorted((b, v) for e in ro.events for b, v in zip(e["samples"], e[k]))
    assert flat(capped, "eps") == flat(whole, "eps") and flat(capped, "rec") == flat(whole, "rec")


@pytest.mark.parametrize("solver,variable", [("euler", True), ("euler", False), ("midpoint", True)])
def test_streaming_session_equals_the_one_shot_rollout(solver, variable):
    """StreamingOdeSession (push each observation, then one non-destructive predict) executes the same operations, in the same
    order and with the same noise, as integrate_latents over the whole history -- oracle standing in for the CUDA engine.
    Covers an observation closer than delta_t to its predecessor (no step before its jump), a target at the last observation's
    own time (served from the current state) and the +-delta_t/2 window; predict leaves the session where it was."""
    from streamingflow_b200.layers.temporal_ode_bayes import NNFOwithBayesianJumps
    from streamingflow_b200.streaming import StreamingOdeSession

    C, h = 8, 6
    m = NNFOwithBayesianJumps(C, C, make_cfg(C, solver=solver, variable=variable)).eval().double()
    m.load_state_dict(so.recipe_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 21, 1.5, torch.float64), strict=True)
    m.__dict__["_engine_factory"] = lambda sd, H, W, n, prec, dev: OracleBackend(sd, H, W, n, prec, dev, torch.float64)
    times = [-0.5, -0.35, -0.32, -0.1, 0.0]
    targets = [0.0, 0.1, 0.25, 0.26, 0.5]
    g = torch.Generator().manual_seed(5)
    hx = torch.tanh(torch.randn(len(times), C, h, h, generator=g, dtype=torch.float64))
    tape = torch.randn(200, C, h, h, generator=g, dtype=torch.float64)
    m._draw_noise = lambda n, hh, ww, device: tape[:max(n, 1)].clone()
    with torch.no_grad():
        final, want = m.integrate_latents(hx, [len(times)], [times], [targets], 0.05)
    used = {"off": 0}

    def sequential(n, hh, ww, device):
        a = used["off"]
        used["off"] += n
        return tape[a:a + max(n, 1)].clone()

    m._draw_noise = sequential
    with torch.no_grad():
        sess = StreamingOdeSession(m, 1, h, h, 0.05, torch.device("cpu"))
        for k, t in enumerate(times):
            sess.push([t], hx[k:k + 1])
        at_obs, clock = sess.state().clone(), list(sess.now)
        got = sess.predict([targets])
        assert torch.equal(sess.state(), at_obs) and sess.now == clock               # the look-ahead did not move the session
        assert (got - want).abs().max() < 1e-12
        assert used["off"] == m.last_rollout.n_eps                                   # same number of noise draws, same order
        again = sess.predict([[0.5]])
        assert again.shape == (1, 1, C, h, h) and torch.isfinite(again).all() and torch.equal(sess.state(), at_obs)
        with pytest.raises(ValueError):
            sess.push([-1.0], hx[:1])                                                # older than the session time
    # two samples with different clocks advance together
    with torch.no_grad():
        s2 = StreamingOdeSession(m, 2, h, h, 0.05, torch.device("cpu"))
        s2.push([-0.5, -0.4], hx[:2])
        s2.push([-0.2, -0.4 + 0.03], hx[2:4])
        out = s2.predict([[0.0, 0.5], [0.0, 0.5]])
    assert out.shape == (2, 2, C, h, h) and torch.isfinite(out).all() and s2.now[1] == -0.4


def _tiny_module(z, dtype):
    from streamingflow_b200.models.future_prediction_ode import FuturePredictionODE

    C = int(z["C"])
    m = FuturePredictionODE(C, C, 4, make_cfg(C)).eval().to(dtype)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(so.recipe_state_dict(shapes, int(z["seed"]), float(z["gain"]), dtype), strict=True)
    m.gru_ode.__dict__["_engine_factory"] = lambda sd, H, W, n, prec, dev: OracleBackend(sd, H, W, n, prec, dev, dtype)
    return m


def test_module_glue_reproduces_reference_output(golden_dir, monkeypatch):
    """FuturePredictionODE.forward of the product (batched rollout, host schedule, pre-drawn noise, window selection,
    torch encoder / decoder / refinement) with the oracle standing in for the CUDA engine == the reference's output."""
    z = np.load(os.path.join(golden_dir, "tiny_full_c8.npz"))
    C, H, B, seed = int(z["C"]), int(z["H"]), int(z["B"]), int(z["seed"])
    dtype = torch.float64
    m = _tiny_module(z, dtype)
    tape = [so.recipe_array(f"eps{i}", (C, H // 4, H // 4), seed, torch.float32) for i in range(64)]
    drawn = {}

    def fake_noise(n, h, w, device):       # the fixture's noise tape stands in for torch's RNG, in slot order
        drawn["n"] = n
        return torch.stack(tape[:n]).to(torch.float32)

    monkeypatch.setattr(m.gru_ode, "_draw_noise", fake_noise)
    # fp32 eps tape (as the engine consumes) vs the fp64 reference tape differ by rounding: regenerate the tape in fp64
    m.gru_ode._draw_noise = lambda n, h, w, device: (drawn.__setitem__("n", n) or torch.stack(
        [so.recipe_array(f"eps{i}", (C, H // 4, H // 4), seed, torch.float64) for i in range(n)]))
    cam = so.recipe_array("cam", (B, 3, C, H, H), seed, dtype)
    lid = so.recipe_array("lidar", (B, 5, C, H, H), seed, dtype)
    with torch.no_grad():
        x, aux = m(torch.zeros(B, 1, C, H, H, dtype=dtype), cam, lid, torch.from_numpy(z["camera_timestamp"]),
                   torch.from_numpy(z["lidar_timestamp"]), torch.from_numpy(z["target_timestamp"]))
    assert aux == 0 and drawn["n"] == int(z["n_eps"])
    ref = torch.from_numpy(z["x_f64"])
    assert ((x - ref).abs().max() / ref.abs().max()).item() < 1e-11


def test_module_requires_cuda_and_has_no_fallback():
    from streamingflow_b200._lib import SfError
    from streamingflow_b200.models.future_prediction_ode import FuturePredictionODE

    m = FuturePredictionODE(64, 64, 4, make_cfg(64)).eval()
    B, H = 1, 8
    with pytest.raises(SfError):
        m(torch.zeros(B, 1, 64, H, H), torch.zeros(B, 1, 64, H, H), None, torch.zeros(B, 1, dtype=torch.float64), None,
          torch.zeros(B, 1, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        m.gru_ode.p_model(torch.zeros(1, 64, 2, 2))


def test_c_abi_library_exports_every_declared_symbol():
    from streamingflow_b200 import _lib

    header = open(os.path.join(ROOT, "include", "sf_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*|const sf_event\*|const int32_t\*)\s+(sf_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert os.path.exists(_lib.LIB_PATH), "libsf_b200.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.sf_abi_version.restype = ctypes.c_int
    assert lib.sf_abi_version() == _lib.SF_ABI_VERSION


def test_packed_stage_plans_c128():
    """The 128-channel stage split (one gate pair / proposal / 128 output channels per launch) replayed on the host."""
    from streamingflow_b200 import engine as en
    import torch.nn.functional as F

    torch.manual_seed(1)
    C = 128
    sd = so.recipe_state_dict(nnfo_shapes(C), 6, 1.0, torch.float32)
    H, W = 4, 5
    rnd = lambda c: torch.randn(1, c, H, W)
    q = lambda t: t.to(torch.bfloat16).float()
    x, s, g1, a, b, t2 = (rnd(C) for _ in range(6))
    y1, q1 = rnd(2 * C), rnd(C)
    src = {-1: x, -2: s, -3: s, en.BUF_G1: g1, en.BUF_G2: g1, en.BUF_A: a, en.BUF_B: b, en.BUF_HH: a, en.BUF_T1: g1, en.BUF_T2: t2,
           en.BUF_Q1: q1, en.BUF_Y1: y1, en.BUF_Q3: y1, en.BUF_Y2: y1}
    conv = lambda inp, w, pad: F.conv2d(q(inp).double(), q(w).double(), None, padding=pad)[0]
    cell = {d.name: d for d in en.cell_stage_defs(sd, "gru_c")}
    assert list(cell) == list(en.CELL_STAGE_NAMES[128])
    xs = torch.cat([x, s], 1)
    acc = en.emulate_stage(cell["gates_1"], False, src)
    ref = torch.cat([conv(xs, sd["gru_c.conv_update_1.weight"], 1), conv(xs, sd["gru_c.conv_reset_1.weight"], 1)])
    assert (acc - ref).abs().max() < 1e-9
    acc = en.emulate_stage(cell["propose_1"], False, src)
    assert (acc[:C] - conv(torch.cat([x, g1], 1), sd["gru_c.conv_state_tilde_1.weight"], 1)).abs().max() < 1e-9
    acc = en.emulate_stage(cell["mix"], False, src)
    ref = torch.cat([conv(t2, sd["gru_c.trusting_gate.0.layers.6.weight"], 1),
                     conv(torch.cat([a, b], 1), sd["gru_c.trusting_gate.0.projection.0.weight"], 0)])
    assert (acc - ref).abs().max() < 1e-9
    prior = {d.name: d for d in en.prior_stage_defs(sd, "p_model") if not isinstance(d, str)}
    assert list(prior) == ["q1", "q2a", "q2b", "q3a", "q3b", "q4a", "q4b", "q5"]
    w2, _ = en._bn_fold(sd, "p_model.model.0.layers.conv_2")
    for h, name in enumerate(("q2a", "q2b")):
        acc = en.emulate_stage(prior[name], False, src)
        ref = torch.cat([conv(q1, w2[128 * h:128 * h + 128], 1), conv(s, sd["p_model.model.0.projection.weight"][128 * h:128 * h + 128], 0)])
        assert (acc - ref).abs().max() < 1e-9 and prior[name].io_off == [128 * h]
    acc = en.emulate_stage(prior["q5"], False, src)
    assert (acc - conv(y1, sd["p_model.model.4.conv.weight"], 1)).abs().max() < 1e-9


def _interpret_codec_graph(graph, bufs_spec, x_in, in_key, out_key, dims):
    """Host interpreter of a codec op list: replays each stage's packed GEMM plan (engine.emulate_stage) and applies the
    epilogue's arithmetic in torch, so the stage graph / weight packing of codec_engine.py is pinned on CPU."""
    from streamingflow_b200 import _lib as L, engine as en
    import torch.nn.functional as F

    bufs = {lvl: {b: torch.zeros(1, ch, *dims[lvl], dtype=torch.float64) for b, ch in chans.items()} for lvl, chans in bufs_spec.items()}
    bufs[in_key[0]][in_key[1]] = x_in.double()
    out32 = None
    for op in graph:
        if op[0] == "pool":
            bufs[op[2][0]][op[2][1]] = F.max_pool2d(bufs[op[1][0]][op[1][1]], 2, 2)
            continue
        if op[0] == "up":
            bufs[op[2][0]][op[2][1]] = F.interpolate(bufs[op[1][0]][op[1][1]], scale_factor=2, mode="nearest")
            continue
        _, lvl, sdef = op
        acc = en.emulate_stage(sdef, True, {b: t.float() for b, t in bufs[lvl].items()})      # split-bf16 replay ~ fp32 accurate
        vec = sdef.vec.double()
        n = max(ck[2].shape[0] for ck in sdef.chunks if ck[3] == 0)
        if sdef.epilogue == L.EPI_BIAS_LRELU:
            v = acc[:n] + vec[:n, None, None]
            act = (sdef.flags >> 1) & 7
            v = F.leaky_relu(v, 0.1) if act == 0 else torch.tanh(v) if act == 1 else v
            dst, off = sdef.io[0], sdef.io_off[0]
        elif sdef.epilogue == L.EPI_RES_ID:
            v = F.leaky_relu(acc[:n] + vec[:n, None, None], 0.1) + bufs[lvl][sdef.io[0]][0, sdef.io_off[0]:sdef.io_off[0] + n]
            dst, off = sdef.io[1], sdef.io_off[1]
        else:
            assert sdef.epilogue == L.EPI_RES_PROJ
            v = F.leaky_relu(acc[:n] + vec[:n, None, None], 0.1) + acc[n:2 * n] + vec[n:2 * n, None, None]
            dst, off = sdef.io[0], sdef.io_off[0]
        bufs[lvl][dst][0, off:off + n] = v
    return bufs[out_key[0]][out_key[1]]


@pytest.mark.parametrize("C,nf", [(64, 64), (128, 64), (128, 128)])
def test_codec_stage_graphs_reproduce_encoder_and_decoder(C, nf):
    """SmallEncoder / SmallDecoder as conv-stage graphs (BatchNorm folded, ConvTranspose as flipped conv, 128-channel output
    groups, residual / projection epilogues) == the oracle's encoder / decoder; input = latent width 64 and 128 (config 5)."""
    from streamingflow_b200 import codec_engine as ce

    sd = so.recipe_state_dict(nnfo_shapes(C, nf), 9, 1.0, torch.float32)
    H, W = 16, 24
    dims = {"A": (H, W), "B": (H // 2, W // 2), "C": (H // 4, W // 4)}
    x = so.recipe_array("bev", (1, C, H, W), 9, torch.float32)
    sd64 = {"g." + k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        want = so.small_encoder(sd64, "g.srvp_encoder", x.double())
        got = _interpret_codec_graph(ce.encoder_graph(sd), ce.enc_bufs(C, nf), x, ce.ENC_IN, ce.ENC_OUT, dims)
    assert got.shape == want.shape and ((got - want).abs().max() / want.abs().max()).item() < 2e-4
    z = torch.tanh(so.recipe_array("lat", (1, C, H // 4, W // 4), 9, torch.float32))
    with torch.no_grad():
        want = so.small_decoder(sd64, "g.srvp_decoder", z.double())
        got = _interpret_codec_graph(ce.decoder_graph(sd), ce.dec_bufs(C, nf), z, ce.DEC_IN, ce.DEC_OUT, dims)
    assert got.shape == want.shape and ((got - want).abs().max() / want.abs().max()).item() < 2e-4


@pytest.mark.parametrize("C, fuse_pw", [(64, False), (64, True), (128, False)])
def test_refinement_stage_graph_reproduces_reference_refinement(C, fuse_pw):
    """SpatialGRU x2 + ConvNeXt Block + DeepLabHead as conv-stage graphs (refine_engine.refine_graph), interpreted on the host,
    == the oracle's refinement (future_prediction_ode.py:56-62); in_channels 64 and 128; fuse_pw: the block's pointwise pair as
    ONE stage (pwconv2 read back from the K-chunked rows appended to the packed matrix, as the kernel's back-to-back GEMM reads them)."""
    from streamingflow_b200 import _lib as L, engine as en, refine_engine as rf
    from streamingflow_b200.models.future_prediction_ode import FuturePredictionODE
    import torch.nn.functional as F

    m = FuturePredictionODE(C, C, 4, make_cfg(C)).eval()
    sd = so.recipe_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 12, 1.0)
    g = rf.refine_graph(sd, fuse_pw=fuse_pw)
    assert len(g["block"]["stages"]) == (1 if fuse_pw else 4 * C // 128 + 1)
    B, T, H, W = (1, 2, 48, 40) if C == 64 else (1, 2, 40, 24)
    x = so.recipe_array("x", (B, T, C, H, W), 12)
    emu = lambda sdef, src: en.emulate_stage(sdef, True, {k: v.float() for k, v in src.items()})
    sig = torch.sigmoid

    def run_gru(i, frames):          # frames [T, C, H, W]
        gi = g[f"gru{i}"]
        state = x[0, 0].double()
        outs = []
        for t in range(T):
            xt = frames[t][None].double()
            acc = emu(gi["gates"], {rf.R_S: state[None], -1: xt})
            vec = gi["gates"].vec.double()
            u = sig(acc[:C] + vec[:C, None, None])
            r = sig(acc[C:2 * C] + vec[C:, None, None])
            acc = emu(gi["propose"], {-1: xt, rf.R_G: ((1 - r) * state)[None]})
            state = (1 - u) * state + u * (acc[:C] + gi["propose"].vec.double()[:, None, None])
            outs.append(emu(gi["dec"], {-1: state[None]})[:C])
        return torch.stack(outs)

    def run_stage(sdef, bufs, img_bias=None):
        acc = emu(sdef, {k: v[None] for k, v in bufs.items()})
        if sdef.flags & L.FLAG_PW_B2B:
            vec = sdef.vec.double()
            t = F.gelu(acc[:4 * C] + vec[:4 * C, None, None])
            rows = en.pack_stage(sdef, True)[1][-2 * 4 * C:].double()                       # [hi rows | lo rows], each [kc][n][64 k]
            w2 = (rows[:4 * C] + rows[4 * C:]).view(4 * C // 64, C, 64).permute(1, 0, 2).reshape(C, 4 * C)
            v = torch.einsum("nk,khw->nhw", w2, t) + vec[4 * C:, None, None] + bufs[sdef.io[0]]
            return sdef.io[1], 0, v
        n = max(ck[2].shape[0] for ck in sdef.chunks if ck[3] == 0)
        v = acc[:n] + sdef.vec.double()[:n, None, None]
        if img_bias is not None:
            v = v + img_bias[:, None, None]
        act = (sdef.flags >> 1) & 7
        v = F.gelu(v) if act == 4 else torch.relu(v) if act == 2 else v
        if sdef.epilogue == L.EPI_RES_ID:
            v = v + bufs[sdef.io[0]][sdef.io_off[0]:sdef.io_off[0] + n]
            return sdef.io[1], sdef.io_off[1], v
        return sdef.io[0], sdef.io_off[0], v

    o0 = run_gru(0, x[0])
    blk = g["block"]
    out_frames = []
    o1_in = []
    for t in range(T):
        dw = F.conv2d(o0[t][None], blk["dw_w"].double(), blk["dw_b"].double(), padding=3, groups=C)[0]
        dw = F.layer_norm(dw.permute(1, 2, 0), (C,), blk["ln_w"].double(), blk["ln_b"].double(), 1e-6).permute(2, 0, 1)
        bufs = {rf.R_O0: o0[t], rf.R_DW: dw, rf.R_P1: torch.zeros(4 * C, H, W, dtype=torch.float64)}
        for sdef in blk["stages"]:
            dst, off, v = run_stage(sdef, bufs)
            if dst == rf.R_P1:
                bufs[rf.R_P1][off:off + v.shape[0]] = v
            else:
                bufs[dst] = v
        o1_in.append(bufs[rf.R_BK])
    o1 = run_gru(1, torch.stack(o1_in))
    pl = g["pool"]
    for t in range(T):
        mean = o1[t].mean(dim=(1, 2))
        vb = torch.relu(pl["pool_w"].double() @ mean + pl["pool_b"].double())
        img_bias = pl["proj_w"].double() @ vb + pl["proj_b"].double()
        bufs = {rf.R_O1: o1[t]}
        for sdef in g["deeplab"]:
            dst, off, v = run_stage(sdef, bufs, img_bias if (sdef.flags & L.FLAG_IMG_BIAS) else None)
            bufs[dst] = v
        out_frames.append(bufs[rf.R_OUT])
    got = torch.stack(out_frames)[None]
    sd64 = {k: v.double() if v.is_floating_point() else v for k, v in sd.items()}
    with torch.no_grad():
        y = so.spatial_gru(sd64, "spatial_grus.0", x.double(), x[:, 0].double())
        y = so.convnext_block(sd64, "res_blocks.0.0", y.reshape(B * T, C, H, W)).view(B, T, C, H, W)
        y = so.spatial_gru(sd64, "spatial_grus.1", y, x[:, 0].double())
        want = so.deeplab_head(sd64, "res_blocks.1", y.reshape(B * T, C, H, W)).view(B, T, C, H, W)
    assert got.shape == want.shape
    assert ((got - want).abs().max() / want.abs().max()).item() < 5e-4


def test_module_can_be_deep_copied_and_pickled(golden_dir):
    """EMA copies (copy.deepcopy) and whole-module saves (pickle / torch.save) must work before AND after the
    first forward; the copy owns its cells (inner API routed to the copy, not the original) and rebuilds its engines."""
    import copy
    import io
    import pickle

    z = np.load(os.path.join(golden_dir, "tiny_full_c8.npz"))
    C, H = int(z["C"]), int(z["H"])
    m = _tiny_module(z, torch.float64)
    c0 = copy.deepcopy(m)
    assert c0.gru_ode.gru_c._owner() is c0.gru_ode and c0.gru_ode.gru_obs.gru_d._owner() is c0.gru_ode
    assert m.gru_ode.gru_c._owner() is m.gru_ode
    args = (torch.zeros(1, 1, C, H, H, dtype=torch.float64), so.recipe_array("cam", (1, 3, C, H, H), 1, torch.float64), None,
            torch.tensor([[-1.0, -0.5, 0.0]], dtype=torch.float64), None, torch.tensor([[0.5, 1.0]], dtype=torch.float64))
    with torch.no_grad():
        torch.manual_seed(5)
        want, _ = m(*args)
    assert m.gru_ode._engines                       # the first forward built an engine (here: the checker backend)
    c1 = copy.deepcopy(m)                           # ... which a copy must not share
    assert c1.gru_ode._engines == {} and c1._refiners == {} and c1.gru_ode.gru_c._owner() is c1.gru_ode
    m.gru_ode.__dict__["_engine_factory"] = None    # lambdas do not pickle; the product never sets a factory
    blob = pickle.dumps(m)
    buf = io.BytesIO()
    torch.save(m, buf)
    for c in (pickle.loads(blob), torch.load(io.BytesIO(buf.getvalue()), weights_only=False)):
        assert c.gru_ode._engines == {} and c.gru_ode.gru_c._owner() is c.gru_ode
        assert all(torch.equal(a, b) for a, b in zip(c.state_dict().values(), m.state_dict().values()))
    c1.gru_ode.__dict__["_engine_factory"] = lambda sd, h, w, n, prec, dev: OracleBackend(sd, h, w, n, prec, dev, torch.float64)
    with torch.no_grad():
        torch.manual_seed(5)
        got, _ = c1(*args)
        dh = c1.gru_ode.gru_c(torch.zeros(1, C, 4, 4, dtype=torch.float64), torch.zeros(1, C, 4, 4, dtype=torch.float64))
    assert torch.equal(got, want) and dh.shape == (1, C, 4, 4)


def test_module_refuses_to_cut_gradients(golden_dir):
    """ADVICE r1 / SURVEY 7.3: the engine has no backward; with autograd recording and anything that requires grad in reach the
    module raises instead of returning tensors without a grad_fn."""
    from streamingflow_b200._lib import SfError

    z = np.load(os.path.join(golden_dir, "tiny_full_c8.npz"))
    C, H = int(z["C"]), int(z["H"])
    m = _tiny_module(z, torch.float64)
    cam = so.recipe_array("cam", (1, 3, C, H, H), 1, torch.float64)
    args = lambda c: (torch.zeros(1, 1, C, H, H, dtype=torch.float64), c, None, torch.tensor([[-1.0, -0.5, 0.0]], dtype=torch.float64), None,
                      torch.tensor([[0.5]], dtype=torch.float64))
    with pytest.raises(SfError, match="no_grad"):
        m(*args(cam))                                            # parameters require grad, autograd is recording
    for p in m.parameters():
        p.requires_grad_(False)
    with pytest.raises(SfError, match="no_grad"):
        m(*args(cam.clone().requires_grad_(True)))               # frozen module, but the input wants a gradient
    x, _ = m(*args(cam))                                         # frozen module, plain input: fine without no_grad
    assert x.shape == (1, 1, C, H, H)
    with pytest.raises(SfError):
        m.gru_ode.infer_state(torch.zeros(1, C, 4, 4, dtype=torch.float64, requires_grad=True))
    with torch.no_grad():
        empty, aux = m(torch.zeros(0, 1, C, H, H), cam[:0], None, torch.zeros(0, 3, dtype=torch.float64), None, torch.zeros(0, 2, dtype=torch.float64))
    assert empty.shape == (0, 2, C, H, H) and aux == 0


def test_float32_stamps_follow_the_reference_float32_schedule(golden_dir):
    """FuturePredictionODE.forward / NNFOwithBayesianJumps.forward pass the stamp tensors' dtype to the scheduler: float32 stamps
    give the float32-arithmetic schedule of the reference (tests/golden/sched_f32.json), not the float64 one."""
    z = np.load(os.path.join(golden_dir, "tiny_full_c8.npz"))
    C, H = int(z["C"]), int(z["H"])
    m = _tiny_module(z, torch.float64)
    cases = json.load(open(os.path.join(golden_dir, "sched_f32.json")))["cases"]
    c = next(c for c in cases if c["variable"] and
             [e.kind for e in so.build_schedule(c["times"], c["targets"], 0.05, True).events] != c["kinds"])
    obs = so.recipe_array("obs", (1, len(c["times"]), C, H, H), 2, torch.float64)
    with torch.no_grad():
        m.gru_ode(torch.tensor(c["times"], dtype=torch.float32), torch.zeros(1, 1, C, H, H, dtype=torch.float64), obs, 0.05,
                  torch.tensor(c["targets"], dtype=torch.float32))
    ro = m.gru_ode.last_rollout
    assert ro.n_state_steps == sum(k == "step" for k in c["kinds"]) and ro.n_jumps == sum(k == "jump" for k in c["kinds"])
    with torch.no_grad():
        m.gru_ode(torch.tensor(c["times"], dtype=torch.float64), torch.zeros(1, 1, C, H, H, dtype=torch.float64), obs, 0.05,
                  torch.tensor(c["targets"], dtype=torch.float64))
    assert m.gru_ode.last_rollout.n_state_steps != ro.n_state_steps


def test_decoder_head_stage_graph_reproduces_reference_decoder(golden_dir):
    """The BEV Decoder head as a conv-stage graph (seg_head_engine.decoder_graph: stride-2 convs as phase chunks over the
    space-to-depth buffers, BasicBlocks with fused shortcuts, 1x1 conv hoisted below the bilinear up-sampling), interpreted on
    the host, == the segmentation logits the UNMODIFIED reference Decoder produced (tests/golden/decoder_seg_c64.npz)."""
    from streamingflow_b200 import _lib as L, engine as en, seg_head_engine as sh
    import torch.nn.functional as F

    z = np.load(os.path.join(golden_dir, "decoder_seg_c64.npz"))
    shapes = {k: tuple(int(t) for t in v.split(",") if t) for k, v in zip(z["shapes_keys"], z["shapes_vals"])}
    sd = so.recipe_state_dict(shapes, int(z["seed"]), float(z["gain"]), torch.float32)
    x = so.recipe_array("dec_in", (1, 2, 64, 32, 32), 17, torch.float64)
    want = torch.from_numpy(z["seg_f64"])
    g = sh.decoder_graph(sd, ["segmentation"])
    H = W = 32
    dims = [(H >> i, W >> i) for i in range(4)]
    outs = []
    for f in range(2):
        bufs = [{b: torch.zeros(1, ch, *dims[l], dtype=torch.float64) for b, ch in chans.items()} for l, chans in enumerate(sh.LEVEL_BUFS)]
        bufs[0][sh.X] = x[:, f]
        for op in g:
            if op[0] == "s2d":
                (ls, bs), (ld, bd), ch = op[1], op[2], op[3]
                src = bufs[ls][bs]
                bufs[ld][bd] = torch.cat([src[:, :, py::2, px::2] for py in range(2) for px in range(2)], dim=1)
            elif op[0] == "upadd":
                (ls, bs), (lk, bk), (ld, bd), ch = op[1], op[2], op[3], op[4]
                bufs[ld][bd] = F.interpolate(bufs[ls][bs], scale_factor=2, mode="bilinear", align_corners=False) + bufs[lk][bk]
            elif op[0] == "head":
                _, key, mod, sig = op
                outs.append(F.conv2d(bufs[0][sh.HD], sd[mod + ".3.weight"].double(), sd[mod + ".3.bias"].double()))
            else:
                _, lvl, sdef = op
                assert len(sdef.chunks) * 2 <= 40, (sdef.name, len(sdef.chunks))      # split mode: two kernel chunks per logical chunk
                acc = en.emulate_stage(sdef, True, {b: t.float() for b, t in bufs[lvl].items()})
                n = max(ck[2].shape[0] for ck in sdef.chunks if ck[3] == 0)
                v = acc[:n] + sdef.vec.double()[:n, None, None]
                relu = ((sdef.flags >> 1) & 7) == L.ACT_RELU
                if sdef.epilogue == L.EPI_RES_ID:
                    assert sdef.flags & L.FLAG_ACT_AFTER_RES
                    v = v + bufs[lvl][sdef.io[0]][0, sdef.io_off[0]:sdef.io_off[0] + n]
                    dst, off = sdef.io[1], sdef.io_off[1]
                else:
                    dst, off = sdef.io[0], sdef.io_off[0]
                bufs[lvl][dst][0, off:off + n] = torch.relu(v) if relu else v
    got = torch.stack(outs, dim=1)
    assert got.shape == want.shape
    err = ((got - want).abs().max() / want.abs().max()).item()
    assert err < 2e-4, err           # the split-bf16 replay carries ~2^-16 per conv
    assert torch.equal(got.argmax(2), want.argmax(2))


def test_inner_api_is_registered_as_torch_custom_ops(golden_dir):
    """north star / SURVEY 8b: the entry points are torch.library operators (torch.ops.sf_b200.*) with fake kernels, and the
    nn.Module methods route through them (here with the checker backend standing in for the CUDA engine on CPU)."""
    import streamingflow_b200.ops  # noqa: F401  (registers the operators)

    for name in ("dual_gru_cell", "infer_state", "ode_step", "integrate_latents"):
        assert hasattr(torch.ops.sf_b200, name)
    z = np.load(os.path.join(golden_dir, "tiny_full_c8.npz"))
    C = int(z["C"])
    m = _tiny_module(z, torch.float32).gru_ode               # fp32 checker backend: same dtypes as the CUDA engine returns
    x, s = torch.randn(1, C, 4, 4), torch.randn(1, C, 4, 4)
    calls = []
    orig = m._cell_impl
    m._cell_impl = lambda *a: (calls.append(1) or orig(*a))
    with torch.no_grad():
        dh = m.gru_c(x, s)                                   # module method -> operator -> engine
        dh2 = torch.ops.sf_b200.dual_gru_cell(x, s, m._op_handle, True)
    assert len(calls) == 2 and torch.equal(dh, dh2)
    torch.library.opcheck(torch.ops.sf_b200.dual_gru_cell, (x, s, m._op_handle, True), test_utils=("test_schema", "test_faketensor"))
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():                                   # shapes without running anything (tracing / torch.compile)
        y, params = torch.ops.sf_b200.infer_state(torch.empty(2, C, 4, 4), m._op_handle)
    assert y.shape == (2, C, 4, 4) and params.shape == (2, 2 * C, 4, 4)


@pytest.mark.parametrize("x3", [False, True])
def test_row_paired_taps_of_a_generic_3x3_stage_replay_to_the_same_accumulator(x3, monkeypatch):
    """engine.pair_rows_if_eligible marks 64 -> 64 3x3 stages (decoder / encoder convs, SpatialGRU proposal, decode, q1) for row-paired
    taps; the packed plan replayed the way the kernel runs it ([dy_hi | dy_lo] operands, block 1 folded one row down, a virtual
    row above the image) gives the accumulator of the plain plan; ineligible stages are left alone."""
    from streamingflow_b200 import _lib as L, engine as en

    assert not en.pair_rows_if_eligible(en.StageDef("c", L.EPI_BIAS_LRELU, torch.zeros(64), [3]).add(7, torch.zeros(64, 64, 3, 3), 0, 1)).flags & L.FLAG_PAIR_ROWS
    monkeypatch.setenv("SF_PAIR_3X3", "1")          # off by default (measured slower on B200); the mechanism stays tested
    g = torch.Generator().manual_seed(4)
    wa, wb = torch.randn(64, 64, 3, 3, generator=g) * 0.1, torch.randn(64, 64, 3, 3, generator=g) * 0.1
    src = {7: torch.randn(1, 64, 21, 13, generator=g), 9: torch.randn(1, 128, 21, 13, generator=g)}
    mk = lambda: en.StageDef("conv", L.EPI_RES_ID, torch.zeros(64), [3, 4]).add(7, wa, 0, 1).add(9, wb, 0, 0, c0=64)
    plain, paired = mk(), en.pair_rows_if_eligible(mk())
    assert paired.flags & L.FLAG_PAIR_ROWS and not plain.flags & L.FLAG_PAIR_ROWS
    a, b = en.emulate_stage(plain, x3, src), en.emulate_stage(paired, x3, src)
    assert (a[:64] - b[:64]).abs().max().item() < 1e-9 * a.abs().max().item()
    assert en.pack_stage(paired, x3)[1].shape == en.pack_stage(plain, x3)[1].shape
    # not eligible: 128 output columns, a 1x1 chunk, a dilated tap, the dual proposal, a 128-channel plan
    wide = en.StageDef("w", L.EPI_BIAS_LRELU, torch.zeros(128), [3]).add(7, torch.randn(128, 64, 3, 3), 0, 1)
    one = en.StageDef("o", L.EPI_BIAS_LRELU, torch.zeros(64), [3]).add(7, wa, 0, 1).add(7, torch.randn(64, 64, 1, 1), 0, 0)
    dil = en.StageDef("d", L.EPI_BIAS_LRELU, torch.zeros(64), [3]).add(7, wa, 0, 1, ox=2)
    dual = en.StageDef("p", L.EPI_PROPOSE, torch.zeros(128), [1, 2, 3, 4]).add(7, wa, 0, 1)
    for sdef in (wide, one, dil, dual):
        assert not en.pair_rows_if_eligible(sdef).flags & L.FLAG_PAIR_ROWS
    assert not en.pair_rows_if_eligible(mk(), C_hidden=128).flags & L.FLAG_PAIR_ROWS


def test_live_noise_slots_are_exactly_the_slots_prior_net_evaluations_read():
    """Rollout.live_eps (what sf_normal_fill_slot_list draws): the noise slots of the events that run the prior net -- every slot when
    the dead evaluations are kept, the slots before a step (and inside a midpoint step) otherwise; slot numbering itself never changes."""
    times = sorted([-1.0, -0.5, 0.0, -0.8, -0.6, -0.4, -0.2, 0.0])
    targets = [-1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0]
    for solver in ("euler", "midpoint"):
        plans = [sc.plan_sample(times, targets, 0.05, True, solver), sc.plan_sample([t + 0.01 for t in times], targets, 0.05, True, solver)]
        full = compile_rollout(plans, [0, 8], solver, True, skip_dead_prior=False)
        live = compile_rollout(plans, [0, 8], solver, True, skip_dead_prior=True)
        assert full.n_eps == live.n_eps and full.live_eps == list(range(full.n_eps))
        want = sorted(s for e in live.events if e["run_prior"] for s in e["eps"])
        assert live.live_eps == want and len(want) == live.n_prior_evals < live.n_eps
        # a slot is live iff the NEXT op of its sample is a step (or it sits inside a midpoint step)
        off = 0
        for plan in plans:
            per = 2 if solver == "midpoint" else 1
            for i, op in enumerate(plan.ops):
                n = per if op.kind == sc.STEP else 1
                nxt_step = i + 1 < len(plan.ops) and plan.ops[i + 1].kind == sc.STEP
                for k in range(n):
                    inside_midpoint = op.kind == sc.STEP and per == 2 and k == 0
                    assert ((off + k) in live.live_eps) == (nxt_step or inside_midpoint)
                off += n
    none = compile_rollout([sc.plan_sample(times, targets, 0.05, True)], [0], "euler", False)
    assert none.live_eps == [] and none.n_eps > 0          # IMPUTE off: the reference still draws, nothing reads
