"""Generate tests/golden/* by RUNNING THE UNMODIFIED REFERENCE (builder container only).

TEST INFRASTRUCTURE.  Usage:  python oracle/gen_golden.py  [--out tests/golden]

The reference has no tests or golden vectors for the ODE head, so the pin
for the oracle is the reference itself, run here on CPU from /root/reference.  Weights, inputs and
noise are produced by the name-keyed recipes in ``oracle/sf_oracle.py`` (``recipe_state_dict``,
``recipe_array``), so the fixtures only hold the reference's OUTPUTS plus the recipe parameters;
the same recipes regenerate the operands on the GPU box, where /root/reference does not exist.

Fixtures written:
  sched.json           step schedules traced from the reference's own control flow (variable / fixed step,
                       euler / midpoint, jittered stamps incl. 1-ulp micro-steps and gap<delta_t cases) and
                       the path-selection result per target.
  tiny_full_c8.npz     FuturePredictionODE.forward, C=8, B=2, 16x16 BEV, jittered stamps, fp64 + fp32 runs.
  c64_latent_*.npz     NNFOwithBayesianJumps.forward, C=64, 32x32 BEV (8x8 latent): per-event latent states,
                       selected states, decoded frames; one file per solver / step-mode / impute variant.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import _refimport as ri  # noqa: E402
from oracle import sf_oracle as so  # noqa: E402

CANON_CAM = [-1.0, -0.5, 0.0]
CANON_LIDAR = [-0.8, -0.6, -0.4, -0.2, 0.0]
CANON_TARGETS = [-1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0]


class Tracer:
    """Wraps ode_step / gru_obs of a reference NNFOwithBayesianJumps instance and srvp_decode to expose the
    event sequence, per-event states and which recorded state each target selected."""

    def __init__(self, nnfo):
        self.n = nnfo
        self.events = []          # (kind, dt, t_after)
        self.states = []
        self.selected = None
        self._ode, self._obs, self._dec = nnfo.ode_step, nnfo.gru_obs.forward, nnfo.srvp_decode

    def __enter__(self):
        def ode_step(state, inp, delta_t, current_time):
            out = self._ode(state, inp, delta_t, current_time)
            t_after = out[2].item() if isinstance(out[2], torch.Tensor) else float(out[2])
            self.events.append(("step", float(delta_t), t_after))
            self.states.append(out[0].detach().clone())
            return out

        def gru_obs(state, p, x_obs):
            out = self._obs(state, p, x_obs)
            self.events.append(("jump", 0.0, float("nan")))
            self.states.append(out[0].detach().clone())
            return out

        def srvp_decode(x, skip=None):
            self.selected = x.detach().clone()
            return self._dec(x, skip)

        self.n.ode_step, self.n.gru_obs.forward, self.n.srvp_decode = ode_step, gru_obs, srvp_decode
        return self

    def __exit__(self, *a):
        self.n.ode_step, self.n.gru_obs.forward, self.n.srvp_decode = self._ode, self._obs, self._dec
        return False

    def selected_event_indices(self):
        idx = []
        for t in range(self.selected.shape[1]):
            hits = [i for i, s in enumerate(self.states) if torch.equal(s, self.selected[:, t])]
            idx.append(hits[-1] if hits else -1)   # latest identical state (states never repeat in practice)
        return idx


def build_nnfo(ref, C, solver, variable, impute, seed, gain, dtype):
    m = ref.tob.NNFOwithBayesianJumps(C, C, ri.make_cfg(C, impute=impute, solver=solver, variable=variable)).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = so.recipe_state_dict(shapes, seed=seed, gain=gain, dtype=dtype)
    m.to(dtype).load_state_dict(sd, strict=True)
    return m, shapes


def gen_schedules(ref, out_dir):
    C, HW = 8, 8
    cases = []
    rng = np.random.RandomState(1234)

    def run_case(times, targets, variable, solver, tag):
        m, _ = build_nnfo(ref, C, solver, variable, True, 3, 1.0, torch.float32)
        obs = torch.zeros(1, len(times), C, HW, HW)
        inp = torch.zeros(1, 1, C, HW, HW)
        with torch.no_grad(), Tracer(m) as tr:
            m(times=torch.tensor(times, dtype=torch.float64), input=inp, obs=obs, delta_t=0.05,
              T=torch.tensor(targets, dtype=torch.float64))
        cases.append(dict(tag=tag, times=[float(t) for t in times], targets=[float(t) for t in targets],
                          variable=variable, solver=solver, delta_t=0.05,
                          kinds=[e[0] for e in tr.events], dts=[e[1] for e in tr.events],
                          t_after=[None if e[0] == "jump" else e[2] for e in tr.events],
                          selected=tr.selected_event_indices()))
        return tr

    canon = sorted(CANON_CAM + CANON_LIDAR)
    for variable in (True, False):
        for solver in ("euler", "midpoint"):
            run_case(canon, CANON_TARGETS, variable, solver, f"canon_var{int(variable)}_{solver}")
    run_case(CANON_CAM, [0.5, 1.0, 1.5, 2.0], True, "euler", "config1_cam_only")
    run_case(canon, CANON_TARGETS[:3] + [0.05 * i for i in range(1, 41)], True, "euler", "streaming_40")
    run_case(canon, CANON_TARGETS[:3] + [0.5 * i for i in range(1, 17)], True, "euler", "long_8s")
    # jittered stamps: keep a spread of step counts, every micro-step case and every gap<delta_t case found
    n_micro = n_gap = n_plain = 0
    for trial in range(400):
        cam = [t + rng.uniform(-0.02, 0.02) for t in CANON_CAM]
        cam[-1] -= rng.uniform(0, 0.03)
        lid = [t + rng.uniform(-0.02, 0.02) for t in CANON_LIDAR]
        tg = [t + rng.uniform(-0.02, 0.02) for t in CANON_TARGETS]
        times = sorted(cam + lid)
        sch = so.build_schedule(times, tg, 0.05, True)
        micro = any(e.kind == "step" and e.dt < 1e-9 for e in sch.events)
        gaps = sum(1 for a, b in zip(times[:-1], times[1:]) if b - a < 0.05)
        keep = (micro and n_micro < 6) or (gaps >= 2 and n_gap < 4) or (not micro and n_plain < 6)
        if not keep:
            continue
        n_micro += micro
        n_gap += (gaps >= 2 and not micro)
        n_plain += (not micro)
        run_case(times, tg, True, "euler", f"jitter_{trial}{'_micro' if micro else ''}")
    # a fixed-step case with jitter (many steps, window selection matters)
    cam = [-1.003, -0.512, -0.004]
    run_case(sorted(cam + [-0.79, -0.61, -0.4, -0.21, 0.0]), [-1.0, -0.5, 0.0, 0.26, 0.5, 0.74, 1.0], False, "euler",
             "fixed_jitter")
    with open(os.path.join(out_dir, "sched.json"), "w") as f:
        json.dump(dict(generator="oracle/gen_golden.py::gen_schedules (reference NNFOwithBayesianJumps.forward traced)",
                       cases=cases), f)
    print(f"sched.json: {len(cases)} cases, {n_micro} with micro-steps")


def gen_schedules_f32(ref, out_dir):
    """Step schedules traced from the reference with FLOAT32 timestamp tensors (a caller that does not follow the nuScenes
    loader's float64 stamps): the comparisons and the variable step are float32 arithmetic there, which moves
    micro-steps and gap < delta_t decisions.  Jittered trials, variable and fixed step."""
    C, HW = 8, 8
    cases = []
    rng = np.random.RandomState(4321)
    for trial in range(60):
        cam = [t + rng.uniform(-0.02, 0.02) for t in CANON_CAM]
        cam[-1] -= rng.uniform(0, 0.03)
        lid = [t + rng.uniform(-0.02, 0.02) for t in CANON_LIDAR]
        tg = [t + rng.uniform(-0.02, 0.02) for t in CANON_TARGETS]
        times = sorted(float(np.float32(t)) for t in cam + lid)
        tg = [float(np.float32(t)) for t in tg]
        variable = trial % 4 != 3
        m, _ = build_nnfo(ref, C, "euler", variable, True, 3, 1.0, torch.float32)
        with torch.no_grad(), Tracer(m) as tr:
            m(times=torch.tensor(times, dtype=torch.float32), input=torch.zeros(1, 1, C, HW, HW), obs=torch.zeros(1, len(times), C, HW, HW),
              delta_t=0.05, T=torch.tensor(tg, dtype=torch.float32))
        cases.append(dict(tag=f"f32_{trial}", times=times, targets=tg, variable=variable, solver="euler", delta_t=0.05,
                          kinds=[e[0] for e in tr.events], dts=[e[1] for e in tr.events],
                          t_after=[None if e[0] == "jump" else e[2] for e in tr.events], selected=tr.selected_event_indices()))
    with open(os.path.join(out_dir, "sched_f32.json"), "w") as f:
        json.dump(dict(generator="oracle/gen_golden.py::gen_schedules_f32 (reference forward traced, float32 stamps)", cases=cases), f)
    n64 = 0
    for c in cases:      # how many of these differ from what float64 arithmetic on the same stamps would do
        sch = so.build_schedule(c["times"], c["targets"], 0.05, c["variable"])
        n64 += [e.kind for e in sch.events] != c["kinds"] or [e.dt for e in sch.events if e.kind == "step"] != [d for k, d in zip(c["kinds"], c["dts"]) if k == "step"]
    print(f"sched_f32.json: {len(cases)} cases, {n64} differ from the float64 schedule of the same stamps")


def gen_tiny_full(ref, out_dir):
    C, H, B, seed, gain = 8, 16, 2, 11, 1.0
    ct = torch.tensor([[-1.0, -0.5, 0.0], [-1.013, -0.492, -0.004]], dtype=torch.float64)
    lt = torch.tensor([[-0.8, -0.6, -0.4, -0.2, 0.0], [-0.81, -0.6, -0.418, -0.2, 0.011]], dtype=torch.float64)
    tt = torch.tensor([CANON_TARGETS, [-1.0, -0.5, 0.0, 0.49, 1.0, 1.52, 2.0]], dtype=torch.float64)
    outs = {}
    n_eps = None
    for name, dtype in (("f64", torch.float64), ("f32", torch.float32)):
        m = ref.FuturePredictionODE(C, C, 4, ri.make_cfg(C)).eval()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.to(dtype).load_state_dict(so.recipe_state_dict(shapes, seed, gain, dtype), strict=True)
        cam = so.recipe_array("cam", (B, 3, C, H, H), seed, dtype)
        lid = so.recipe_array("lidar", (B, 5, C, H, H), seed, dtype)
        eps = [so.recipe_array(f"eps{i}", (1, C, H // 4, H // 4), seed, dtype) for i in range(64)]
        with torch.no_grad(), ri.EpsTape(replay=eps) as tape:
            x, aux = m(torch.zeros(B, 1, C, H, H, dtype=dtype), cam, lid, ct, lt, tt)
        assert aux == 0
        n_eps = len(tape.tape)
        if name == "f64":
            outs["x_f64"] = x.numpy()
            x64 = x
        else:
            outs["ref_f32_x_err"] = float((x.double() - x64).abs().max() / x64.abs().max())
        # B>1 == sequential B=1 on one noise stream
        if name == "f64":
            xs = []
            with torch.no_grad(), ri.EpsTape(replay=eps):
                for b in range(B):
                    xs.append(m(torch.zeros(1, 1, C, H, H, dtype=dtype), cam[b:b + 1], lid[b:b + 1], ct[b:b + 1],
                                lt[b:b + 1], tt[b:b + 1])[0])
            assert torch.equal(torch.cat(xs, 0), x), "reference: batched != sequential"
        if name == "f64":
            outs["shapes_keys"] = np.array(list(shapes.keys()))
            outs["shapes_vals"] = np.array([",".join(map(str, v)) for v in shapes.values()])
    np.savez_compressed(os.path.join(out_dir, "tiny_full_c8.npz"), C=C, H=H, B=B, seed=seed, gain=gain, n_eps=n_eps,
                        camera_timestamp=ct.numpy(), lidar_timestamp=lt.numpy(), target_timestamp=tt.numpy(), **outs)
    print("tiny_full_c8.npz: x", outs["x_f64"].shape, "n_eps", n_eps)


def gen_c64_latent(ref, out_dir):
    canon = sorted(CANON_CAM + CANON_LIDAR)
    variants = [
        ("euler_var", "euler", True, True, canon, CANON_TARGETS),
        ("midpoint_var", "midpoint", True, True, canon, [-1.0, 0.0, 0.5, 1.0]),
        ("euler_fixed", "euler", False, True, CANON_CAM, [0.0, 0.25, 0.5]),
        ("euler_var_noimpute", "euler", True, False, canon, [-0.5, 0.0, 1.0, 2.0]),
        ("euler_var_jitter", "euler", True, True,
         sorted([-1.013, -0.492, -0.004, -0.81, -0.6, -0.418, -0.2, 0.011]), [-1.0, -0.5, 0.0, 0.49, 1.0, 1.52, 2.0]),
    ]
    gen_latent(ref, out_dir, 64, 32, 7, variants)


def gen_c128_latent(ref, out_dir):
    """The 128-channel network of BASELINE config 5 (in_channels = latent_dim = OUT_CHANNELS = FILTER_SIZE = 128) on a small grid."""
    canon = sorted(CANON_CAM + CANON_LIDAR)
    gen_latent(ref, out_dir, 128, 16, 11, [("euler_var", "euler", True, True, canon, CANON_TARGETS),
                                           ("midpoint_var", "midpoint", True, True, CANON_CAM, [0.0, 0.5, 1.0])])


def gen_latent(ref, out_dir, C, H, seed, variants, gain=1.0):
    for tag, solver, variable, impute, times, targets in variants:
        res = {}
        for name, dtype in (("f64", torch.float64), ("f32", torch.float32)):
            m, shapes = build_nnfo(ref, C, solver, variable, impute, seed, gain, dtype)
            obs = so.recipe_array("obs", (1, len(times), C, H, H), seed, dtype)
            eps = [so.recipe_array(f"eps{i}", (1, C, H // 4, H // 4), seed, dtype) for i in range(256)]
            with torch.no_grad(), ri.EpsTape(replay=eps) as tape, Tracer(m) as tr:
                state, loss, x = m(times=torch.tensor(times, dtype=torch.float64),
                                   input=torch.zeros(1, 1, C, H, H, dtype=dtype), obs=obs, delta_t=0.05,
                                   T=torch.tensor(targets, dtype=torch.float64))
            assert loss == 0
            if name == "f64":       # ground truth, stored as fp32 (6e-8 relative)
                res["states_f64"] = torch.cat(tr.states, 0).to(torch.float32).numpy()
                res["final_f64"] = state.to(torch.float32).numpy()
                res["x_f64"] = x[:, [0, -1]].to(torch.float32).numpy()
                st64, x64 = torch.cat(tr.states, 0), x
            else:                   # the reference's own fp32 run: keep only its distance to fp64
                res["ref_f32_state_err"] = float((torch.cat(tr.states, 0).double() - st64).abs().max() / st64.abs().max())
                res["ref_f32_x_err"] = float((x.double() - x64).abs().max() / x64.abs().max())
            if name == "f64":
                res["selected"] = np.array(tr.selected_event_indices())
                res["kinds"] = np.array([e[0] for e in tr.events])
                res["dts"] = np.array([e[1] for e in tr.events])
                res["n_eps"] = len(tape.tape)
                res["encoded_f64"] = m.srvp_encoder(obs[0]).detach().to(torch.float32).numpy()
        np.savez_compressed(os.path.join(out_dir, f"c{C}_latent_{tag}.npz"), C=C, H=H, seed=seed, gain=gain, solver=solver,
                            variable=variable, impute=impute, times=np.array(times), targets=np.array(targets), **res)
        print(f"c{C}_latent_{tag}.npz: events {len(res['kinds'])}, n_eps {res['n_eps']}, "
              f"ref f32-vs-f64 state {res['ref_f32_state_err']:.2e} x {res['ref_f32_x_err']:.2e}")


def gen_decoder(ref, out_dir):
    """The reference Decoder's segmentation branch (models/decoder.py) on a small input: pins oracle.seg_decoder and records
    the Decoder's parameter shapes so the GPU box can rebuild recipe weights by name."""
    import importlib

    dec_mod = importlib.import_module("streamingflow.models.decoder")
    gate = dict(perceive_hdmap=False, predict_pedestrian=False, predict_instance=False, predict_future_flow=False, planning=False)
    d = dec_mod.Decoder(in_channels=64, n_classes=2, n_present=3, n_hdmap=2, predict_gate=gate).eval().double()
    shapes = {k: tuple(v.shape) for k, v in d.state_dict().items()}
    d.load_state_dict(so.recipe_state_dict(shapes, 17, 1.0, torch.float64), strict=True)
    x = so.recipe_array("dec_in", (1, 2, 64, 32, 32), 17, torch.float64)
    with torch.no_grad():
        seg = d(x)["segmentation"]
    np.savez_compressed(os.path.join(out_dir, "decoder_seg_c64.npz"), seed=17, gain=1.0, seg_f64=seg.numpy(),
                        shapes_keys=np.array(list(shapes.keys())), shapes_vals=np.array([",".join(map(str, v)) for v in shapes.values()]))
    print("decoder_seg_c64.npz:", tuple(seg.shape), "argmax class-1 fraction", float((seg.argmax(2) == 1).double().mean()))
    # every head (all predict gates on), n_present = 2: pins oracle.bev_decoder
    gate = {k: True for k in gate}
    d = dec_mod.Decoder(in_channels=64, n_classes=2, n_present=2, n_hdmap=2, predict_gate=gate).eval().double()
    shapes = {k: tuple(v.shape) for k, v in d.state_dict().items()}
    d.load_state_dict(so.recipe_state_dict(shapes, 19, 1.0, torch.float64), strict=True)
    x = so.recipe_array("dec_in", (1, 3, 64, 32, 32), 19, torch.float64)
    with torch.no_grad():
        out = d(x)
    np.savez_compressed(os.path.join(out_dir, "decoder_all_c64.npz"), seed=19, gain=1.0, n_present=2, **{k: v.numpy() for k, v in out.items()},
                        shapes_keys=np.array(list(shapes.keys())), shapes_vals=np.array([",".join(map(str, v)) for v in shapes.values()]))
    print("decoder_all_c64.npz:", {k: tuple(v.shape) for k, v in out.items()})


ARGMAX_SEEDS = tuple(range(101, 141))


def gen_argmax(ref, out_dir, H=200, gain=1.15, keep=3):
    """The graded end product (trainer.py:230-231): argmax over the segmentation logits of the reference Decoder applied to
    the reference FuturePredictionODE output, at the BASELINE config-1 shapes (B = 1, 3 camera frames, 4 future targets, BEV
    HxHx64), with 'energised' weights so the masks are non-trivial: the recipe draws N(0, gain^2 / fan_in), torch's default init
    has variance 1 / (3 fan_in), so recipe gain 1.15 corresponds to SURVEY 8c's "default init x 2.0".  For several weight/input seeds
    the unmodified reference is run in fp64; the seeds whose smallest |logit_0 - logit_1| is largest are kept (a mask can only
    be asserted bit-exact when no pixel is a near-tie), their masks are stored bit-packed, with the margin statistics."""
    import importlib

    dec_mod = importlib.import_module("streamingflow.models.decoder")
    gate = dict(perceive_hdmap=False, predict_pedestrian=False, predict_instance=False, predict_future_flow=False, planning=False)
    C = 64
    ct = torch.tensor([CANON_CAM], dtype=torch.float64)
    tt = torch.tensor([[0.5, 1.0, 1.5, 2.0]], dtype=torch.float64)
    rows = []
    for seed in ARGMAX_SEEDS:
        m = ref.FuturePredictionODE(C, C, 4, ri.make_cfg(C)).eval()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.double().load_state_dict(so.recipe_state_dict(shapes, seed, gain, torch.float64), strict=True)
        d = dec_mod.Decoder(in_channels=C, n_classes=2, n_present=3, n_hdmap=2, predict_gate=gate).eval().double()
        dshapes = {k: tuple(v.shape) for k, v in d.state_dict().items()}
        d.load_state_dict(so.recipe_state_dict(dshapes, seed, gain, torch.float64), strict=True)
        cam = so.recipe_array("cam", (1, 3, C, H, H), seed, torch.float64)
        eps = [so.recipe_array(f"eps{i}", (1, C, H // 4, H // 4), seed, torch.float64) for i in range(16)]
        with torch.no_grad(), ri.EpsTape(replay=eps) as tape:
            x, _ = m(torch.zeros(1, 1, C, H, H, dtype=torch.float64), cam, None, ct, None, tt)
            seg = d(x)["segmentation"]
        margin = (seg[:, :, 0] - seg[:, :, 1]).abs()
        mask = seg.argmax(dim=2)
        hist = np.histogram(margin.numpy().ravel(), bins=[0, 1e-4, 1e-3, 1e-2, 3e-2, 1e-1, 3e-1, 1, 1e9])[0].tolist()
        rows.append(dict(seed=seed, margin_hist=hist, min_margin=float(margin.min()), max_logit=float(seg.abs().max()), frac1=float((mask == 1).double().mean()),
                         frac_lt_1e2=float((margin < 1e-2).double().mean()), frac_lt_1e1=float((margin < 1e-1).double().mean()),
                         x_absmax=float(x.abs().max()), n_eps=len(tape.tape), mask=mask.numpy().astype(np.uint8)))
        print({k: v for k, v in rows[-1].items() if k != "mask"}, flush=True)
    # a constant mask would make the comparison trivial: only seeds with at least 20 pixels of the minority class qualify
    minority = lambda r: min(r["frac1"], 1.0 - r["frac1"]) * r["mask"].size
    rows.sort(key=lambda r: (minority(r) < 20, -r["min_margin"] / r["max_logit"]))
    meta = [{k: v for k, v in r.items() if k != "mask"} for r in rows]
    kept = rows[:keep]
    np.savez_compressed(os.path.join(out_dir, "argmax_c64.npz"), H=H, gain=gain, seeds=np.array([r["seed"] for r in kept]),
                        min_margin=np.array([r["min_margin"] for r in kept]), max_logit=np.array([r["max_logit"] for r in kept]),
                        frac1=np.array([r["frac1"] for r in kept]), n_eps=np.array([r["n_eps"] for r in kept]),
                        masks=np.packbits(np.stack([r["mask"] for r in kept]), axis=-1), all_seeds=json.dumps(meta),
                        dshapes_keys=np.array(list(dshapes.keys())), dshapes_vals=np.array([",".join(map(str, v)) for v in dshapes.values()]))
    print("argmax_c64.npz: kept seeds", [r["seed"] for r in kept])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  "tests", "golden"))
    ap.add_argument("--only", default="", help="comma-separated subset: schedules,tiny,c64,c128,decoder,argmax")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    torch.set_num_threads(8)
    ref = ri.import_reference()
    steps = dict(schedules=gen_schedules, schedules_f32=gen_schedules_f32, tiny=gen_tiny_full, c64=gen_c64_latent, c128=gen_c128_latent, decoder=gen_decoder, argmax=gen_argmax)
    for name, fn in steps.items():
        if not args.only or name in args.only.split(","):
            fn(ref, args.out)


if __name__ == "__main__":
    main()
