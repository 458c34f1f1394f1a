"""Push / pull use of the GRU-ODE-Bayes path.

The reference's streaming evaluation (evaluate_streaming.py:119-133) calls the whole model again for every new frame, i.e.
re-integrates the latent state from the first observation each time.  The ODE state is a running quantity, so a session
keeps it on the device instead:

    session = StreamingOdeSession(model.gru_ode, batch=B, h=50, w=50)
    session.push(t_obs, latent_frames)          # advance to the observation (ODE steps), then the Bayesian jump
    future = session.predict(targets)           # roll a COPY of the state forward; the session stays at the observation

Both calls execute exactly the operations the one-shot rollout would execute for the same history
(schedule.plan_sample: the advance-to-observation loop, temporal_ode_bayes.py:539-581, and the advance-to-target loop with
its +-delta_t/2 record window, :585-622), in the same order and with the same double-precision time arithmetic, and draw the
noise in the same order -- so push x n + one predict reproduces NNFOwithBayesianJumps.integrate_latents for a timeline whose
targets lie at or after the last observation (tests/test_host_logic.py, tests/test_gpu_rollout.py).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .rollout import compile_rollout
from .schedule import JUMP, STEP, Op, SamplePlan, _window_pick


class StreamingOdeSession:
    def __init__(self, ode, batch: int, h: int, w: int, delta_t: float = 0.05, device: Optional[torch.device] = None):
        self.ode, self.B, self.h, self.w, self.delta_t = ode, int(batch), int(h), int(w), float(delta_t)
        self.device = torch.device(device) if device is not None else next(ode.parameters()).device
        self.eng = ode._engine_for(self.h, self.w, self.B, self.device)
        self.reset()

    # ------------------------------------------------------------------ state
    def reset(self):
        """Forget the history: zero state (the reference's initial state, temporal_ode_bayes.py:508-511), no current time."""
        self.eng.zero_state(0)
        self.now: List[Optional[float]] = [None] * self.B       # per sample: the loop's current_time (python double, like .item())
        self.last_obs: List[Optional[float]] = [None] * self.B  # per sample: time stamp of the last jump (a recorded state)
        self.n_pushed = 0

    def state(self) -> torch.Tensor:
        """The current latent state [B, C, h, w] (fp32), at the time of the last pushed observation."""
        return self.eng.unpack_f32(self.eng.state32[0], self.B)

    # ------------------------------------------------------------------ one observation per sample
    def push(self, t_obs: Sequence[float], latents: torch.Tensor) -> None:
        """t_obs[b]: time of sample b's new observation (>= its previous one); latents: its encoded frame [B, C, h, w].
        Runs, per sample, the ODE steps from the current time to the observation and then the jump."""
        assert len(t_obs) == self.B and latents.shape[0] == self.B
        plans = []
        for b, t in enumerate(float(t) for t in t_obs):
            now = t if self.now[b] is None else self.now[b]
            if t < now:
                raise ValueError(f"sample {b}: observation at {t} is older than the session time {now}")
            ops: List[Op] = []
            while now <= t - self.delta_t:                                   # schedule.plan_sample, first loop
                hstep = (t - now) if self.ode.use_variable_ode_step else self.delta_t
                now = now + hstep
                ops.append(Op(STEP, hstep, -1, now))
            ops.append(Op(JUMP, 0.0, 0, t))
            self.now[b], self.last_obs[b] = now, t          # a jump does not move current_time (it may trail t_obs by < delta_t)
            plans.append(self._plan(ops, []))
        self._run(plans, latents)
        self.n_pushed += 1

    # ------------------------------------------------------------------ non-destructive look-ahead
    def predict(self, targets: Sequence[Sequence[float]]) -> torch.Tensor:
        """targets[b]: increasing times at or after sample b's current time.  Returns the selected latents [B, T, C, h, w]
        (the reference's +-delta_t/2 window rule); the session's own state, sampled input and time are left untouched."""
        assert len(targets) == self.B and self.n_pushed > 0
        T = len(targets[0])
        half = 0.5 * self.delta_t
        plans, current = [], []
        for b in range(self.B):
            now = self.now[b]
            ops: List[Op] = []
            stamps, stamp_op = [self.last_obs[b]], [-1]                      # -1: the state as it is now (recorded at the last jump)
            for t_goal in (float(t) for t in targets[b]):
                while now < t_goal:                                          # schedule.plan_sample, second loop
                    hstep = (t_goal - now) if self.ode.use_variable_ode_step else self.delta_t
                    now = now + hstep
                    ops.append(Op(STEP, hstep, -1, now))
                    if t_goal - half < now < t_goal + half:
                        stamps.append(now)
                        stamp_op.append(len(ops) - 1)
            picks = [stamp_op[_window_pick(stamps, float(t), half)] for t in targets[b]]
            current.append([i for i, p in enumerate(picks) if p < 0])
            plans.append(self._plan(ops, [p for p in picks if p >= 0]))
        snap = self.eng.snapshot(self.B)
        here = self.state()
        ro = self._run(plans, None)
        flat = [s for slots in ro.out_slots for s in slots]
        rolled = self.eng.unpack_path(flat) if flat else None                # one gather for every recorded state
        out = torch.empty((self.B, T) + tuple(here.shape[1:]), dtype=here.dtype, device=here.device)
        k = 0
        for b in range(self.B):
            for i in range(T):
                if i in current[b]:
                    out[b, i] = here[b]
                else:
                    out[b, i] = rolled[k]
                    k += 1
        self.eng.restore(snap, self.B)
        return out

    # ------------------------------------------------------------------ plumbing
    def _plan(self, ops: List[Op], picks: List[int]) -> SamplePlan:
        per_step = 2 if self.ode.solver == "midpoint" else 1
        return SamplePlan(ops, picks, sum(per_step if o.kind == STEP else 1 for o in ops))

    def _run(self, plans: List[SamplePlan], latents: Optional[torch.Tensor]):
        ode, eng = self.ode, self.eng
        # observation image of sample b = b.  A push leaves the session's sampled input behind for whatever comes next (keep the
        # last op's prior evaluation); a look-ahead is rolled back afterwards, its last input is dead
        ro = compile_rollout(plans, list(range(self.B)), ode.solver, bool(ode.impute), skip_dead_prior=ode.skip_dead_prior,
                             keep_last_input=latents is not None)
        if latents is not None:
            eng.bind_observations(latents)
        eng.ensure_path_slots(max(ro.n_path, 1))
        if ro.n_eps:
            eng.bind_eps(ode._draw_noise(ro.n_eps, self.h, self.w, self.device))
        if ro.events:
            eng.run_rollout(ro.events)
        return ro
