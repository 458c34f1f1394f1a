"""Compiles per-sample step plans into batched engine events.

The reference integrates one sample at a time (future_prediction_ode.py:36-51).  Samples are
independent, so round r of the batched rollout executes the r-th engine event of every sample that still
has one, grouped by (event kind, state buffers, x source): one set of stage launches per group with the
sample list as a device-side index table.  Schedules may differ between samples (jitter, micro-steps),
which only changes group membership -- a sample is never padded with a fake step, because even a dt = 0
step would redraw the sampled input (SURVEY 7.3 'Schedule parity').

Noise slots follow the reference's consumption order exactly: sample-major, then event order, one
standard-normal tensor per infer_state call (two per midpoint step).

Dead prior-net evaluations.  The reference calls ``infer_state`` (p_model + rsample) after EVERY op, but the sampled input it
produces is read only by a following ``ode_step``: an observation jump feeds the cell the observation itself
(``GRUObservationCell.forward(state, p, X_obs)`` ignores ``p``, temporal_ode_bayes.py:327-344) and nothing reads the input after
the last op (:606-624 select and decode STATES).  ``skip_dead_prior`` (default) therefore runs the prior net only where its
sample is consumed -- before a step, and inside a midpoint step -- which leaves every state, hence every output, bit-identical
(the parity fixtures compare per-event states).  The noise slots are still numbered as if every call drew: the RNG stream of
live draws is the reference's.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

from .schedule import JUMP, STEP, SamplePlan

# keep in sync with engine.py buffer ids
BUF_X, BUF_OBS, BUF_ZERO = 2, 3, 4


@dataclass
class Rollout:
    events: List[dict] = field(default_factory=list)      # engine event dicts (see OdeEngine.build_table)
    n_eps: int = 0                                         # noise tensors to draw, in slot order
    n_path: int = 0                                        # recorded states
    out_slots: List[List[int]] = field(default_factory=list)   # per sample, per target: path slot
    n_state_steps: int = 0                                 # ode_step calls over all samples (the metric's unit)
    n_jumps: int = 0
    n_cell_evals: int = 0
    n_prior_evals: int = 0
    trace_slots: List[List[int]] = field(default_factory=list)  # record_all: per sample, per op: path slot
    live_eps: List[int] = field(default_factory=list)      # noise slots some prior-net evaluation reads (all of them unless dead evaluations are skipped)


def compile_rollout(plans: Sequence[SamplePlan], obs_base: Sequence[int], solver: str, impute: bool,
                    record_all: bool = False, obs_index=None, max_group: int = 0, skip_dead_prior: bool = True,
                    keep_last_input: bool = False) -> Rollout:
    """record_all additionally records the state after EVERY op (debug / parity traces).  obs_index(b, k) overrides the
    image index of sample b's k-th observation in the OBS buffer (default: sample-major obs_base[b] + k).
    skip_dead_prior: do not evaluate the prior net after an op whose sampled input nothing reads (module docstring);
    keep_last_input: the input sampled after the LAST op is live (a streaming session continues from it later)."""
    ro = Rollout()
    if obs_index is None:
        obs_index = lambda b, k: obs_base[b] + k
    per_sample: List[List[dict]] = []
    eps = 0
    for b, plan in enumerate(plans):
        picked: Dict[int, int] = {}
        slots = []
        for op_idx in plan.picks:
            if op_idx not in picked:
                picked[op_idx] = ro.n_path
                ro.n_path += 1
            slots.append(picked[op_idx])
        ro.out_slots.append(slots)
        if record_all:
            for i in range(len(plan.ops)):
                if i not in picked:
                    picked[i] = ro.n_path
                    ro.n_path += 1
            ro.trace_slots.append([picked[i] for i in range(len(plan.ops))])
        evs: List[dict] = []
        n_ops = len(plan.ops)
        for i, op in enumerate(plan.ops):
            rec = picked.get(i, -1)
            # is the input sampled after this op read by anything?  only by a following ode_step
            live = (plan.ops[i + 1].kind == STEP) if i + 1 < n_ops else keep_last_input
            prior = bool(impute) and (live or not skip_dead_prior)
            if op.kind == JUMP:
                evs.append(dict(kind=JUMP, x_buf=BUF_OBS, x_img=obs_index(b, op.obs), s_in=0, s_base=0, s_out=0, dt=0.0, eps=eps,
                                rec=rec, run_prior=prior))
                eps += 1
                ro.n_jumps += 1
            else:
                xb, xi = (BUF_X, b) if impute else (BUF_ZERO, 0)
                if solver == "euler":
                    evs.append(dict(kind=STEP, x_buf=xb, x_img=xi, s_in=0, s_base=0, s_out=0, dt=op.dt, eps=eps, rec=rec,
                                    run_prior=prior))
                    eps += 1
                elif solver == "midpoint":
                    # k = s + dt/2 f(x, s); pk = infer(k)   |   s = s + dt f(pk, k); x = infer(s)      (tob:449-454)
                    evs.append(dict(kind=STEP, x_buf=xb, x_img=xi, s_in=0, s_base=0, s_out=1, dt=op.dt / 2, eps=eps, rec=-1,
                                    run_prior=True))
                    evs.append(dict(kind=STEP, x_buf=BUF_X, x_img=b, s_in=1, s_base=0, s_out=0, dt=op.dt, eps=eps + 1, rec=rec,
                                    run_prior=prior))
                    eps += 2
                else:
                    raise ValueError(f"Unknown solver '{solver}'.")
                ro.n_state_steps += 1
        per_sample.append(evs)
    ro.n_eps = eps
    ro.live_eps = sorted(e["eps"] for evs in per_sample for e in evs if e["run_prior"])
    depth = max((len(e) for e in per_sample), default=0)
    for r in range(depth):
        groups: Dict[Tuple, dict] = {}
        for b, evs in enumerate(per_sample):
            if r >= len(evs):
                continue
            e = evs[r]
            key = (e["kind"], e["x_buf"], e["s_in"], e["s_base"], e["s_out"], e["run_prior"])
            g = groups.setdefault(key, dict(kind=e["kind"], x_buf=e["x_buf"], s_in=e["s_in"], s_base=e["s_base"], s_out=e["s_out"],
                                            run_cell=1, run_prior=int(e["run_prior"]), want_f32=0, samples=[], x_img=[], rec=[],
                                            eps=[], dt=[]))
            g["samples"].append(b)
            g["x_img"].append(e["x_img"])
            g["rec"].append(e["rec"])
            g["eps"].append(e["eps"])
            g["dt"].append(e["dt"])
        for g in groups.values():
            ro.n_cell_evals += len(g["samples"])
            ro.n_prior_evals += len(g["samples"]) * g["run_prior"]
            n = len(g["samples"])
            step = max_group if max_group and max_group > 0 else n
            for i in range(0, n, step):          # optional cap on the samples per event (keeps an event's intermediates in L2)
                part = dict(g)
                for k in ("samples", "x_img", "rec", "eps", "dt"):
                    part[k] = g[k][i:i + step]
                ro.events.append(part)
    return ro
