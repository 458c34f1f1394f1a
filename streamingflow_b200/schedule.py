"""Host-side step scheduling for the GRU-ODE-Bayes rollout (pure Python doubles, no device syncs).

Mirrors the control flow of ``NNFOwithBayesianJumps.forward`` in the reference
(streamingflow/layers/temporal_ode_bayes.py:508 start time, :539-553 advance-to-observation loop,
:562-581 jump and record, :585-604 advance-to-target loop with the +-delta_t/2 record window,
:606-622 output selection).  The reference evaluates this with ``.item()`` round trips between every
kernel; here the whole (sample -> op list) table is known before the first launch.

Floating-point fidelity matters: ``t += (t_next - t)`` may land one ulp short of
``t_next``, which makes the reference take an extra ~1e-16 s step (one more cell evaluation and one
more noise draw).  Python floats are IEEE doubles, the same arithmetic ``.item()`` yields, so the
quirk is reproduced by construction; tests pin it against traces of the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

STEP, JUMP = 0, 1


@dataclass(frozen=True)
class Op:
    kind: int        # STEP | JUMP
    dt: float        # step size (double, as the reference multiplies it); 0.0 for a jump
    obs: int         # observation index (jump) or -1
    t: float         # time after the op


@dataclass
class SamplePlan:
    ops: List[Op]
    picks: List[int]          # per target: index into ops whose resulting state is the output frame
    n_noise: int              # standard-normal tensors the rollout consumes (solver dependent)

    @property
    def n_steps(self) -> int:
        return sum(1 for o in self.ops if o.kind == STEP)


def _window_pick(stamps: Sequence[float], when: float, half: float) -> int:
    """Latest recorded stamp strictly inside (when-half, when+half); else the nearest one (first on ties),
    as ``np.argmin`` does in the reference (:612-620)."""
    inside = [i for i, s in enumerate(stamps) if when - half < s < when + half]
    if inside:
        return inside[-1]
    best, best_d = 0, None
    for i, s in enumerate(stamps):
        d = abs(s - when)
        if best_d is None or d < best_d:
            best, best_d = i, d
    return best


def plan_sample(obs_times: Sequence[float], targets: Sequence[float], delta_t: float, variable_step: bool,
                solver: str = "euler", obs_dtype: str = "float64", target_dtype: str = "float64") -> SamplePlan:
    """obs_times must already be in processing order (sorted by the caller exactly like the reference's dict sort).

    obs_dtype / target_dtype: dtype of the timestamp TENSORS the caller passed.  The reference keeps ``current_time`` as a
    python double but compares and subtracts it against 0-dim tensors of the stamps' own dtype (:540-545, :586-590), so with
    float32 stamps the comparisons, the variable step ``t_next - current_time`` and ``current_time += dt`` are float32
    arithmetic; a fixed step (``current_time += delta_t``, two python floats) stays double.  The selection (:606-620) runs on
    ``.item()`` doubles in both cases.  numpy scalars of the stamp dtype reproduce this op for op (np.float64 == python float)."""
    if len(obs_times) == 0:
        raise ValueError("at least one observation is required (the reference takes times.min())")
    import numpy as np

    D1 = np.float32 if obs_dtype == "float32" else np.float64
    D2 = np.float32 if target_dtype == "float32" else np.float64
    obs_times = [float(t) for t in obs_times]
    targets = [float(t) for t in targets]
    now = min(obs_times)              # times.min().item(): a python double
    ops: List[Op] = []
    stamps: List[float] = []      # recorded times
    stamp_op: List[int] = []      # op index that produced each recorded state
    half = 0.5 * delta_t
    for k, t_obs in enumerate(obs_times):
        t1 = D1(t_obs)
        while D1(now) <= t1 - D1(delta_t):
            if variable_step:
                h = t1 - D1(now)
                now = float(D1(now) + h)
                h = float(h)
            else:
                h = delta_t
                now = now + h
            ops.append(Op(STEP, h, -1, now))
        ops.append(Op(JUMP, 0.0, k, t_obs))
        stamps.append(t_obs)
        stamp_op.append(len(ops) - 1)
    for t_goal in targets:
        t2 = D2(t_goal)
        while D2(now) < t2:
            if variable_step:
                h = t2 - D2(now)
                now = float(D2(now) + h)
                h = float(h)
            else:
                h = delta_t
                now = now + h
            ops.append(Op(STEP, h, -1, now))
            if t2 - D2(half) < D2(now) < t2 + D2(half):
                stamps.append(now)
                stamp_op.append(len(ops) - 1)
    picks = [stamp_op[_window_pick(stamps, t, half)] for t in targets]
    per_step = 2 if solver == "midpoint" else 1
    n_noise = sum(per_step if o.kind == STEP else 1 for o in ops)
    return SamplePlan(ops, picks, n_noise)


def merge_observations(camera_t: Sequence[float], lidar_t: Sequence[float] | None) -> List[Tuple[float, int, int]]:
    """Processing order of one sample's observations: the reference fills a dict with the camera frames, then the
    lidar frames, and sorts it by time with a stable sort (future_prediction_ode.py:37-45), so camera wins ties and
    duplicates are all kept.  Returns [(time, sensor, index)] with sensor 0 = camera, 1 = lidar."""
    items = [(float(t), 0, i) for i, t in enumerate(camera_t)]
    if lidar_t is not None:
        items += [(float(t), 1, i) for i, t in enumerate(lidar_t)]
    items.sort(key=lambda it: it[0])
    return items
