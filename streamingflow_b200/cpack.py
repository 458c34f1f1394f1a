"""Python view of the library's weight packers and host schedule (include/sf_b200.h: sf_pack_cell_weights,
sf_pack_pmodel_weights, sf_rollout_plan_*; streamingflow_b200/csrc/sf_ode.cu).

The engine packs the ODE head's weights through these calls, so the Python host and a C host ("C host") feed
the kernels the same bytes; ``engine.cell_stage_defs`` / ``prior_stage_defs`` / ``pack_stage`` remain as the generic packer of
the codec / refinement / decoder-head stages and as the cross-check of this one (tests/test_host_logic.py).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L


@dataclass
class PackedItem:
    name: str
    se_layer: int = -1                     # >= 0: squeeze-excite marker (fc1 / fc2 only)
    epilogue: int = 0
    flags: int = 0
    chunks: List[dict] = field(default_factory=list)
    w: Optional[torch.Tensor] = None       # bf16 [rows, 64]
    vec: Optional[torch.Tensor] = None
    io: List[int] = field(default_factory=list)
    io_off: List[int] = field(default_factory=list)
    fold_se: Optional[int] = None
    w32: Optional[torch.Tensor] = None
    row_meta: Optional[torch.Tensor] = None
    fc1: Optional[torch.Tensor] = None
    fc2: Optional[torch.Tensor] = None


def tensor_table(sd: Dict[str, torch.Tensor], prefix: str = ""):
    """(sf_tensor array, keep-alive list) of the floating-point entries of ``sd`` whose key starts with ``prefix``."""
    keep, rows = [], []
    for k, v in sd.items():
        if not k.startswith(prefix) or not torch.is_tensor(v) or not v.is_floating_point():
            continue
        t = v.detach().to("cpu", torch.float32).contiguous()
        name = k.encode()
        keep += [t, name]
        rows.append(L.Tensor(name, t.data_ptr(), t.numel()))
    arr = (L.Tensor * max(1, len(rows)))(*rows)
    return arr, len(rows), keep


def _f32(ptr, n):
    return torch.from_numpy(np.ctypeslib.as_array(ptr, shape=(n,)).copy()) if n else torch.zeros(0)


def _unpack(lib, handle) -> List[PackedItem]:
    out = []
    try:
        n = L.check(lib.sf_packed_count(handle), "sf_packed_count")
        for i in range(n):
            d = L.StageDesc()
            L.check(lib.sf_packed_get(handle, i, C.byref(d)), "sf_packed_get")
            it = PackedItem(name=d.name.decode(), se_layer=d.se_layer)
            if d.se_layer >= 0:
                it.fc1, it.fc2 = _f32(d.fc1, d.n_fc), _f32(d.fc2, d.n_fc)
                out.append(it)
                continue
            it.epilogue, it.flags = d.epilogue, d.flags
            it.chunks = [{f: getattr(d.chunks[c], f) for f, _ in L.Chunk._fields_} for c in range(d.n_chunks)]
            raw = np.ctypeslib.as_array(C.cast(d.w, C.POINTER(C.c_uint16)), shape=(d.w_rows * 64,)).copy()
            it.w = torch.from_numpy(raw.view(np.int16)).view(torch.bfloat16).view(d.w_rows, 64)
            it.vec = _f32(d.vec, d.n_vec)
            it.io = [d.io[k] for k in range(d.n_io)]
            it.io_off = [d.io_off[k] for k in range(d.n_io)]
            if d.fold_se >= 0:
                it.fold_se = d.fold_se
                it.w32 = _f32(d.w32, d.w_rows * 64).view(d.w_rows, 64)
                it.row_meta = torch.from_numpy(np.ctypeslib.as_array(d.row_meta, shape=(d.w_rows,)).copy())
            out.append(it)
    finally:
        lib.sf_packed_free(handle)
    return out


def pack_cell(sd: Dict[str, torch.Tensor], prefix: str, x3: bool, pair_rows: bool = True, b2b: bool = True, pair3: bool = False) -> List[PackedItem]:
    """One dual-GRU cell (``prefix`` = '...gru_c.' or '...gru_obs.gru_d.') packed by the library."""
    lib = L.load()
    arr, n, keep = tensor_table(sd, prefix)
    h = C.c_void_p()
    opts = (L.PACK_PAIR_ROWS if pair_rows else 0) | (L.PACK_B2B if b2b else 0) | (L.PACK_PAIR_3X3 if pair3 else 0)
    L.check(lib.sf_pack_cell_weights(arr, n, prefix.encode(), L.PREC_BF16X3 if x3 else L.PREC_BF16, opts, C.byref(h)), "sf_pack_cell_weights")
    del keep
    return _unpack(lib, h)


def pack_pmodel(sd: Dict[str, torch.Tensor], prefix: str, x3: bool, fold_se: bool, pair3: bool = False) -> List[PackedItem]:
    """p_model (``prefix`` = '...p_model.') packed by the library: q1 .. q5 with the two SE markers in launch order."""
    lib = L.load()
    arr, n, keep = tensor_table(sd, prefix)
    h = C.c_void_p()
    L.check(lib.sf_pack_pmodel_weights(arr, n, prefix.encode(), L.PREC_BF16X3 if x3 else L.PREC_BF16,
                                       (L.PACK_FOLD_SE if fold_se else 0) | (L.PACK_PAIR_3X3 if pair3 else 0), C.byref(h)),
            "sf_pack_pmodel_weights")
    del keep
    return _unpack(lib, h)


@dataclass
class CRollout:
    events: List[L.Event]
    table: np.ndarray
    out_slots: List[List[int]]
    info: Dict[str, int]


def plan_rollout(obs_times: Sequence[Sequence[float]], targets: Sequence[Sequence[float]], delta_t: float, variable_step: bool,
                 solver: str = "euler", impute: bool = True, obs_dtype: str = "float64", target_dtype: str = "float64",
                 all_prior: bool = False, keep_last_input: bool = False) -> CRollout:
    """sf_rollout_plan_create: the library's own host schedule (the C restatement of schedule.plan_sample + rollout.compile_rollout
    for samples with equally many observations, obs image index b * n_obs + k)."""
    lib = L.load()
    B, n_obs, n_t = len(obs_times), len(obs_times[0]), len(targets[0])
    assert all(len(o) == n_obs for o in obs_times) and all(len(t) == n_t for t in targets) and len(targets) == B
    ob = (C.c_double * (B * n_obs))(*[float(t) for row in obs_times for t in row])
    tg = (C.c_double * (B * n_t))(*[float(t) for row in targets for t in row])
    h = C.c_void_p()
    if solver not in ("euler", "midpoint"):
        raise ValueError(f"Unknown solver '{solver}'.")
    L.check(lib.sf_rollout_plan_create(ob, n_obs, tg, n_t, B, float(delta_t), int(variable_step), 0 if solver == "euler" else 1, int(impute),
                                       int(obs_dtype == "float32"), int(target_dtype == "float32"), int(all_prior) | (2 if keep_last_input else 0),
                                       C.byref(h)), "sf_rollout_plan_create")
    try:
        info = L.RolloutInfo()
        L.check(lib.sf_rollout_plan_info(h, C.byref(info)), "sf_rollout_plan_info")
        evp, tbp, osp = lib.sf_rollout_plan_events(h), lib.sf_rollout_plan_table(h), lib.sf_rollout_plan_out_slots(h)
        events = [L.Event(*[getattr(evp[i], f) for f, _ in L.Event._fields_]) for i in range(info.n_events)]
        table = np.ctypeslib.as_array(tbp, shape=(info.n_table,)).copy() if info.n_table else np.zeros(0, np.int32)
        slots = [[osp[b * n_t + j] for j in range(n_t)] for b in range(B)]
        return CRollout(events, table, slots, {f: getattr(info, f) for f, _ in L.RolloutInfo._fields_})
    finally:
        lib.sf_rollout_plan_free(h)
