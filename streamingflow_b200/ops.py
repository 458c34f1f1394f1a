"""The path's entry points as torch custom ops (``torch.ops.sf_b200.*``), so that the dispatcher, ``torch.compile`` and graph
capture see them as opaque operators (north star: "a thin C-ABI torch custom-op extension").

The arithmetic lives behind the C ABI (include/sf_b200.h -> libsf_b200.so, bound with ctypes in _lib.py); an op only routes a
call to the engine of the ``NNFOwithBayesianJumps`` instance registered under ``handle`` (engines own device workspaces and
packed weights, which a stateless operator signature cannot carry).  The nn.Module methods of the reference's inner API
(``gru_c(x, state)``, ``gru_obs(state, p, X_obs)``, ``infer_state``, ``ode_step``, the latent rollout) call these ops.

    sf_b200::dual_gru_cell(x, state, handle, derivative)      -> dh | new state                 temporal_ode_bayes.py:92-131, 239-275
    sf_b200::infer_state(state, handle)                        -> (sample, params)               :463-477
    sf_b200::ode_step(state, input, dt, handle)                -> (state, input)                 :436-459
    sf_b200::integrate_latents(hx_obs, handle, plan)           -> (final states, selected latents)   :507-622

Ops draw their noise from the module's noise source (the global CUDA generator in the reference's order), i.e. they are
impure like the reference's ``rsample``; fake (meta) kernels give shapes for tracing.
"""
from __future__ import annotations

import weakref
from typing import Dict, Tuple

import torch
from torch.library import custom_op

_MODULES: Dict[int, "weakref.ref"] = {}
_PLANS: Dict[int, tuple] = {}
_next_plan = [1]


def register_module(m) -> int:
    h = id(m)
    _MODULES[h] = weakref.ref(m, lambda _r, _h=h: _MODULES.pop(_h, None))
    return h


def _module(handle: int):
    ref = _MODULES.get(handle)
    m = ref() if ref is not None else None
    if m is None:
        raise RuntimeError(f"sf_b200: no live NNFOwithBayesianJumps registered under handle {handle}")
    return m


def stash_plan(plan: tuple) -> int:
    """Host-side arguments of a rollout (observation counts, times, targets, delta_t, stamp dtypes) do not fit an operator
    signature of tensors and scalars: the caller parks them here and passes the key."""
    k = _next_plan[0]
    _next_plan[0] += 1
    _PLANS[k] = plan
    return k


@custom_op("sf_b200::dual_gru_cell", mutates_args=())
def dual_gru_cell(x: torch.Tensor, state: torch.Tensor, handle: int, derivative: bool) -> torch.Tensor:
    return _module(handle)._cell_impl(x, state, derivative)


@dual_gru_cell.register_fake
def _(x, state, handle, derivative):
    return torch.empty_like(state, dtype=torch.float32)


@custom_op("sf_b200::infer_state", mutates_args=())
def infer_state(state: torch.Tensor, handle: int) -> Tuple[torch.Tensor, torch.Tensor]:
    return _module(handle)._infer_state_impl(state)


@infer_state.register_fake
def _(state, handle):
    n, c, h, w = state.shape
    return state.new_empty((n, c, h, w), dtype=torch.float32), state.new_empty((n, 2 * c, h, w), dtype=torch.float32)


@custom_op("sf_b200::ode_step", mutates_args=())
def ode_step(state: torch.Tensor, input: torch.Tensor, dt: float, handle: int) -> Tuple[torch.Tensor, torch.Tensor]:
    return _module(handle)._ode_step_impl(state, input, dt)


@ode_step.register_fake
def _(state, input, dt, handle):
    return torch.empty_like(state, dtype=torch.float32), torch.empty_like(state, dtype=torch.float32)


@custom_op("sf_b200::integrate_latents", mutates_args=())
def integrate_latents(hx_obs: torch.Tensor, handle: int, plan: int) -> Tuple[torch.Tensor, torch.Tensor]:
    obs_counts, times, targets, delta_t, stamp_dtypes = _PLANS.pop(plan)
    return _module(handle)._integrate_impl(hx_obs, obs_counts, times, targets, delta_t, stamp_dtypes=stamp_dtypes)


@integrate_latents.register_fake
def _(hx_obs, handle, plan):
    obs_counts, times, targets, delta_t, stamp_dtypes = _PLANS[plan]
    _, c, h, w = hx_obs.shape
    B, T = len(obs_counts), len(targets[0])
    return hx_obs.new_empty((B, c, h, w), dtype=torch.float32), hx_obs.new_empty((B, T, c, h, w), dtype=torch.float32)
