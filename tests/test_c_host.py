"""CPU tests of the library's host layer (streamingflow_b200/csrc/sf_ode.cu; include/sf_b200.h "The ODE head driven from this
header alone"): the C weight packers against the torch restatement in engine.py, the C host schedule against schedule.py /
rollout.py and against the reference's traced schedules, and the size / error conventions of sf_ode_*.  No CUDA compute."""
import ctypes as C
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import sf_oracle as so
from streamingflow_b200 import _lib as L
from streamingflow_b200 import cpack
from streamingflow_b200 import engine as en
from streamingflow_b200.config import ode_cfg
from streamingflow_b200.engine import OdeEngine
from streamingflow_b200.layers.temporal_ode_bayes import NNFOwithBayesianJumps
from streamingflow_b200.rollout import compile_rollout
from streamingflow_b200.schedule import plan_sample


def _bn_fold_ieee(sd, p):
    """engine._bn_fold in correctly rounded fp32 (numpy).  torch's CPU sqrt goes through MKL VML, which is NOT correctly rounded
    (sqrt(1.1747403f) comes out one ulp low), so the torch restatement differs from the library -- and from torch's own CUDA
    sqrt -- in the last bit of a few folded scales; the library follows IEEE."""
    f = lambda k: sd[p + k].float().numpy()
    scale = f(".norm.weight") / np.sqrt(f(".norm.running_var") + np.float32(1e-5))
    bias = f(".norm.bias") - f(".norm.running_mean") * scale
    return torch.from_numpy(f(".conv.weight") * scale[:, None, None, None]), torch.from_numpy(bias)


def _weights(Cc, seed=5):
    m = NNFOwithBayesianJumps(Cc, Cc, ode_cfg(Cc)).eval()
    return so.recipe_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed, 1.0)


def _same_stage(g, d, x3):
    chunks, wp = en.pack_stage(d, x3)
    assert (g.name, g.epilogue, g.flags, g.io, g.io_off) == (d.name, d.epilogue, d.flags, d.io, d.io_off), d.name
    assert torch.equal(g.vec, d.vec.float()), d.name
    assert g.chunks == chunks, d.name
    assert g.w.shape == wp.shape and torch.equal(g.w.view(torch.int16), wp.view(torch.int16)), d.name


@pytest.mark.parametrize("Cc", [64, 128])
def test_c_weight_packers_equal_the_torch_restatement_bit_for_bit(Cc, monkeypatch):
    """sf_pack_cell_weights / sf_pack_pmodel_weights (what the engine and a C host feed the kernels) == cell_stage_defs /
    prior_stage_defs + pack_stage + pack_stage_master: stage list, epilogues, flags, io, chunk tables, constant vectors and every
    byte of the packed bf16 matrices (cat[state, state] fold, BatchNorm fold, tap order, row pairing, fused 1x1, hi / lo split)."""
    monkeypatch.setattr(en, "_bn_fold", _bn_fold_ieee)
    sd = _weights(Cc)
    options = ((True, True, True), (False, True, False), (True, False, True), (False, False, False), (True, True, False)) if Cc == 64 else ((True, True, True),)
    for x3 in (False, True):
        for pair, b2b, pair3 in options:
            for pre in ("gru_c", "gru_obs.gru_d"):
                want = en.cell_stage_defs(sd, pre, pair, b2b, pair3)
                got = cpack.pack_cell(sd, pre + ".", x3, pair, b2b, pair3)
                if Cc == 64:
                    assert bool(next(g for g in got if g.name == "decode").flags & L.FLAG_PAIR_ROWS) == pair3
                assert [d.name for d in want] == [g.name for g in got]
                for d, g in zip(want, got):
                    _same_stage(g, d, x3)
        for fold, pair3 in ((False, True), (True, True), (True, False)):
            want = en.prior_stage_defs(sd, "p_model", fold_se=fold, pair3=pair3)
            got = cpack.pack_pmodel(sd, "p_model.", x3, fold, pair3)
            assert len(want) == len(got)
            for d, g in zip(want, got):
                if isinstance(d, str):
                    idx = 1 if g.se_layer == 0 else 3
                    assert g.name == d and g.se_layer == int(d[2])
                    assert torch.equal(g.fc1, sd[f"p_model.model.{idx}.fc.0.weight"].float().reshape(-1))
                    assert torch.equal(g.fc2, sd[f"p_model.model.{idx}.fc.2.weight"].float().reshape(-1))
                    continue
                _same_stage(g, d, x3)
                assert g.fold_se == d.fold_se, d.name
                if d.fold_se is not None:
                    w32, meta = en.pack_stage_master(d, x3)
                    assert torch.equal(g.w32, w32) and torch.equal(g.row_meta, meta), d.name


def test_c_packer_takes_a_prefix_and_reports_what_is_missing():
    sd = {"gru_ode." + k: v for k, v in _weights(64).items()}
    got = cpack.pack_cell(sd, "gru_ode.gru_c.", False)
    assert [g.name for g in got] == ["gates", "propose", "decode", "trunk", "mix"]
    del sd["gru_ode.gru_c.conv_reset_2.weight"]
    with pytest.raises(L.SfError, match="conv_reset_2.weight"):
        cpack.pack_cell(sd, "gru_ode.gru_c.", False)
    bad = {k: (v[:32] if k.endswith("conv_decoder_2.bias") else v) for k, v in _weights(64).items()}
    with pytest.raises(L.SfError, match="64 or 128"):
        cpack.pack_cell(bad, "gru_c.", False)


def _python_rollout(obs, tg, dt, var, solver, impute, od, td, all_prior, keep_last):
    plans = [plan_sample(o, t, dt, var, solver, od, td) for o, t in zip(obs, tg)]
    n_obs = len(obs[0])
    ro = compile_rollout(plans, [b * n_obs for b in range(len(obs))], solver, impute, skip_dead_prior=not all_prior, keep_last_input=keep_last)
    table, evs = OdeEngine.build_table(ro.events)
    return ro, table, evs


def test_c_host_schedule_equals_the_python_schedule():
    """sf_rollout_plan_create == plan_sample + compile_rollout + build_table: event structs, the int32 event table (sample ids, x
    images, record slots, noise slots, float32 dt bits), output slots and the counters -- jittered stamps, divergent schedules
    inside a batch, euler / midpoint, variable / fixed step, IMPUTE off, float32 stamp arithmetic, dead-prior and keep-last flags."""
    rng = random.Random(0)
    base = sorted([-1.0, -0.5, 0.0] + [-0.8, -0.6, -0.4, -0.2, 0.0])
    tgt = [-1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0]
    for trial in range(120):
        B = rng.choice([1, 2, 3, 8])
        jit = (lambda: rng.uniform(-0.02, 0.02)) if trial % 3 else (lambda: 0.0)
        obs = [sorted(t + jit() for t in base) for _ in range(B)]
        tg = [sorted(t + jit() for t in tgt) for _ in range(B)]
        for var in (True, False):
            for solver in ("euler", "midpoint"):
                for od, td in (("float64", "float64"), ("float32", "float32"), ("float32", "float64")):
                    args = (obs, tg, 0.05, var, solver, trial % 5 != 0, od, td, trial % 7 == 0, trial % 11 == 0)
                    ro, table, evs = _python_rollout(*args)
                    c = cpack.plan_rollout(*args)
                    assert np.array_equal(table, c.table)
                    assert len(evs) == len(c.events)
                    for a, b in zip(evs, c.events):
                        assert all(getattr(a, f) == getattr(b, f) for f, _ in L.Event._fields_)
                    assert ro.out_slots == c.out_slots
                    i = c.info
                    assert (i["n_eps"], i["n_path"], i["n_state_steps"], i["n_jumps"], i["n_cell_evals"], i["n_prior_evals"]) == \
                        (ro.n_eps, ro.n_path, ro.n_state_steps, ro.n_jumps, ro.n_cell_evals, ro.n_prior_evals)


@pytest.mark.parametrize("name, dtype", [("sched.json", "float64"), ("sched_f32.json", "float32")])
def test_c_host_schedule_matches_reference_traces(golden_dir, name, dtype):
    """The C schedule against the op sequences traced from the UNMODIFIED reference (tests/golden/sched*.json): op kinds, step
    sizes (the table carries them as float32), and the selected output states, incl. the 1-ulp micro-step cases."""
    for c in json.load(open(os.path.join(golden_dir, name)))["cases"]:
        r = cpack.plan_rollout([c["times"]], [c["targets"]], c["delta_t"], c["variable"], c["solver"], True, dtype, dtype, all_prior=True)
        kinds, dts, rec_of_op = [], [], []
        per_op = 2 if c["solver"] == "midpoint" else 1
        pending = 0
        for ev in r.events:
            assert ev.n_active == 1
            rows = r.table[ev.table_off:ev.table_off + 5]
            dt = float(rows[4:5].view(np.float32)[0])
            if ev.kind == 1:
                kinds.append("jump"); dts.append(0.0); rec_of_op.append(int(rows[2]))
            else:
                pending += 1
                if pending == per_op:        # a midpoint step = two events; the second one carries the full dt and the record slot
                    kinds.append("step"); dts.append(dt); rec_of_op.append(int(rows[2]))
                    pending = 0
        assert kinds == c["kinds"], c["tag"]
        for k, got, want in zip(kinds, dts, c["dts"]):
            if k == "step":
                assert got == float(np.float32(want)), (c["tag"], got, want)
        selected = [rec_of_op.index(s) for s in r.out_slots[0]]
        assert selected == c["selected"], c["tag"]
        steps = kinds.count("step")
        assert r.info["n_state_steps"] == steps and r.info["n_jumps"] == len(kinds) - steps
        assert r.info["n_eps"] == steps * per_op + r.info["n_jumps"]


def test_c_host_schedule_error_conventions():
    lib = L.load()
    h = C.c_void_p()
    ob, tg = (C.c_double * 1)(0.0), (C.c_double * 1)(1.0)
    assert lib.sf_rollout_plan_create(ob, 1, tg, 1, 1, 0.05, 1, 2, 1, 0, 0, 0, C.byref(h)) == -1           # solver check, tob:386
    assert b"solver" in lib.sf_last_error()
    assert lib.sf_rollout_plan_create(ob, 0, tg, 1, 1, 0.05, 1, 0, 1, 0, 0, 0, C.byref(h)) == -1           # no observation: times.min() of nothing
    with pytest.raises(ValueError):
        cpack.plan_rollout([[0.0]], [[1.0]], 0.05, True, solver="rk4")


@pytest.mark.parametrize("Cc, x3", [(64, False), (64, True), (128, False)])
def test_ode_workspace_query_is_pure_host_and_accounts_for_every_buffer(Cc, x3):
    """sf_ode_query_workspace needs no device: the size covers the activation planes, the fp32 masters, the path / noise / observation
    tensors and the packed weights, and grows with each of them."""
    lib = L.load()

    def query(B=2, H=24, W=20, path=5, obs=7, eps=11, opts=L.PACK_DEFAULT):
        g = L.Geometry(B, H, W, Cc, L.PREC_BF16X3 if x3 else L.PREC_BF16, 0)
        o = L.OdeOptions(path, obs, eps, opts)
        n = C.c_size_t()
        L.check(lib.sf_ode_query_workspace(C.byref(g), C.byref(o), C.byref(n)), "sf_ode_query_workspace")
        return n.value

    base = query()
    hw, planes = 24 * 20, 2 if x3 else 1
    # lower bound from the header's buffer list: 13 C-wide + 3 2C-wide per-sample activation buffers (SE outputs folded away),
    # 4 fp32 state-sized masters + x32 + params32, path / eps / obs
    act = 2 * hw * Cc * 2 * planes * (13 + 2 * 3)
    f32 = 2 * hw * Cc * 4 * (5 + 2)
    assert base >= act + f32 + (5 + 11) * hw * Cc * 4 + 7 * hw * Cc * 2 * planes
    assert query(path=6) - base == hw * Cc * 4 or query(path=6) - base == ((hw * Cc * 4 + 255) // 256) * 256
    assert query(eps=12) > base and query(obs=8) > base and query(B=3) > base
    # unfolded SE layers: two more 2C-wide activation buffers instead of the per-sample scaled copies of q3 / q5's weights
    unfolded = query(opts=L.PACK_PAIR_ROWS | L.PACK_B2B)
    assert unfolded != base and unfolded - query(opts=L.PACK_PAIR_ROWS | L.PACK_B2B, B=1) > 2 * hw * 2 * Cc * 2 * planes
    g = L.Geometry(2, 24, 20, 96, 0, 0)
    n = C.c_size_t()
    assert lib.sf_ode_query_workspace(C.byref(g), None, C.byref(n)) == -1 and b"64 or 128" in lib.sf_last_error()


def test_ode_create_fails_loudly_without_a_b200():
    """sf_ode_create on a host without the GPU: the weights pack (pure host), then the plan refuses -- a negative status and a message,
    no silent CPU path (the size query and the packers are the only entry points that work here)."""
    if torch.cuda.is_available():
        pytest.skip("needs a host without a GPU")
    lib = L.load()
    sd = _weights(64)
    hot = {k: v for k, v in sd.items() if k.startswith(("gru_c.", "gru_obs.", "p_model."))}
    arr, n, keep = cpack.tensor_table(hot)
    g = L.Geometry(1, 12, 12, 64, L.PREC_BF16, 0)
    o = L.OdeOptions(4, 3, 9, L.PACK_DEFAULT)
    h = C.c_void_p()
    fake_ws = C.c_void_p(1 << 20)          # never dereferenced: creation fails before the first device call on it
    rc = lib.sf_ode_create(C.byref(g), C.byref(o), arr, n, b"", fake_ws, C.c_size_t(1 << 40), C.byref(h))
    assert rc < 0 and not h.value and len(lib.sf_last_error()) > 0
    bad = lib.sf_ode_create(C.byref(g), C.byref(o), arr, n, b"", C.c_void_p((1 << 20) + 8), C.c_size_t(1 << 40), C.byref(h))
    assert bad == -1 and b"aligned" in lib.sf_last_error()
    assert lib.sf_ode_create(C.byref(g), C.byref(o), arr, n, b"nope.", fake_ws, C.c_size_t(1 << 40), C.byref(h)) == -1      # prefix strips every tensor


def test_c_merge_of_camera_and_lidar_stamps_equals_the_reference_order():
    """sf_merge_observations == schedule.merge_observations (the reference's dict fill + stable sort, future_prediction_ode.py:37-45):
    camera wins ties, duplicates stay, LiDAR may be absent."""
    from streamingflow_b200.schedule import merge_observations

    lib = L.load()
    rng = random.Random(5)
    cases = [([-1.0, -0.5, 0.0], [-0.8, -0.5, 0.0, 0.0]), ([-1.0, -0.5, 0.0], None), ([0.0], [0.0, 0.0])]
    for _ in range(50):
        cases.append(([round(rng.uniform(-1, 0), 1) for _ in range(rng.randint(1, 4))], [round(rng.uniform(-1, 0), 1) for _ in range(rng.randint(0, 6))]))
    for cam, lid in cases:
        n_l = len(lid) if lid else 0
        ct = (C.c_double * len(cam))(*cam)
        lt = (C.c_double * max(1, n_l))(*(lid or [0.0]))
        times = (C.c_double * (len(cam) + n_l))()
        src = (C.c_int32 * (len(cam) + n_l))()
        n = lib.sf_merge_observations(ct, len(cam), lt if n_l else None, n_l, times, src)
        want = merge_observations(cam, lid if n_l else None)
        assert n == len(want)
        assert [(times[i], src[i] >> 16, src[i] & 0xFFFF) for i in range(n)] == [(t, s, i) for t, s, i in want]
