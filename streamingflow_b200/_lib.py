"""ctypes binding of libsf_b200.so (the C ABI in include/sf_b200.h).

There is no fallback: if the shared library is missing or the device is not a B200, every entry point
raises.  Build it with ``python -m streamingflow_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsf_b200.so")

SF_ABI_VERSION = 3
PREC_BF16, PREC_BF16X3 = 0, 1
(EPI_GATES, EPI_PROPOSE, EPI_DECODE, EPI_LNGELU, EPI_MIX, EPI_BIAS_LRELU, EPI_RES_PROJ, EPI_RES_ID, EPI_SAMPLE) = range(9)
(F32_STATE0, F32_STATE1, F32_A, F32_B, F32_PATH, F32_SE_SUMS, F32_EPS, F32_X, F32_PARAMS, F32_ERRFLAG, F32_OUT, F32_IMG_BIAS) = range(12)
ACT_LRELU, ACT_TANH, ACT_RELU, ACT_NONE, ACT_GELU = 0, 1, 2, 3, 4
FLAG_KEEP_A32, FLAG_OUT32, FLAG_SINGLE, FLAG_IMG_BIAS, FLAG_RES_SE_SCALE = 1, 16, 32, 64, 128   # bit 8: which SE layer scales the residual
FLAG_PAIR_ROWS = 512          # lngelu stages at C = 64: vertically adjacent taps paired into one MMA of twice the width
FLAG_ACT_AFTER_RES = 2048     # res_id: out = act(conv + bias + residual) (ResNet BasicBlock) instead of act(conv + bias) + residual
FLAG_DERIV = 4096             # propose: output u (s~ - s) (GRU-ODE derivative) instead of the blended state
FLAG_B2B = 1024               # lngelu stages at C = 64: a 1x1 conv + LN + GELU fused behind the stage (weights appended to w_packed)
FLAG_PW_B2B = 8192            # res_id at C = 64 (bf16): pwconv1 -> GELU -> pwconv2 -> + residual of the ConvNeXt block in ONE launch
SRC_X, SRC_STATE_IN, SRC_STATE_OUT = -1, -2, -3
SE_MAX_PARTIALS = 160        # SF_SE_MAX_PARTIALS in include/sf_b200.h
SE_ITEM_BASE = 1000          # event-graph item: SE reduce + apply (activation pass)
SE_FOLD_ITEM_BASE = 2000     # event-graph item: SE reduce + scales folded into the consumers' weights


class Chunk(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("buf", "plane", "c0", "R", "n", "nrep", "col", "wrow", "init", "ox", "oy")]


class Geometry(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("max_images", "H", "W", "C", "precision", "device")]


class Event(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("kind", "n_active", "x_buf", "s_in", "s_base", "s_out", "run_cell", "run_prior",
                                         "want_f32", "table_off")]


class Tensor(C.Structure):
    """sf_tensor: a named fp32 tensor in host memory (reference state_dict key, torch layout)."""
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


class StageDesc(C.Structure):
    """sf_stage_desc: read-only view of one item of a packed stage list (sf_packed_get)."""
    _fields_ = [("name", C.c_char_p), ("se_layer", C.c_int32), ("epilogue", C.c_int32), ("flags", C.c_int32), ("n_chunks", C.c_int32),
                ("chunks", C.POINTER(Chunk)), ("w", C.c_void_p), ("w_rows", C.c_int32), ("vec", C.POINTER(C.c_float)), ("n_vec", C.c_int32),
                ("n_io", C.c_int32), ("io", C.POINTER(C.c_int32)), ("io_off", C.POINTER(C.c_int32)), ("fold_se", C.c_int32),
                ("w32", C.POINTER(C.c_float)), ("row_meta", C.POINTER(C.c_int32)), ("fc1", C.POINTER(C.c_float)), ("fc2", C.POINTER(C.c_float)), ("n_fc", C.c_int32)]


class OdeOptions(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("path_slots", "obs_images", "eps_slots", "pack_options")]


class RolloutInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_events", "n_table", "n_eps", "n_path", "n_state_steps", "n_jumps", "n_cell_evals", "n_prior_evals")]


PACK_PAIR_ROWS, PACK_B2B, PACK_FOLD_SE, PACK_PAIR_3X3 = 1, 2, 4, 8
PACK_DEFAULT = PACK_PAIR_ROWS | PACK_B2B | PACK_FOLD_SE
(ODE_OBS_HI, ODE_OBS_LO, ODE_EPS, ODE_PATH, ODE_STATE0, ODE_STATE1, ODE_X32, ODE_PARAMS32, ODE_ERRFLAG) = range(9)

EXPORTS = {
    "sf_abi_version": (C.c_int, []),
    "sf_last_error": (C.c_char_p, []),
    "sf_device_supported": (C.c_int, [C.c_int]),
    "sf_plan_create": (C.c_int, [C.POINTER(Geometry), C.POINTER(C.c_void_p)]),
    "sf_plan_destroy": (C.c_int, [C.c_void_p]),
    "sf_plan_bind_act": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "sf_plan_bind_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "sf_plan_define_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Chunk), C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, C.c_int]),
    "sf_plan_define_stage_fold": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sf_plan_define_se": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "sf_plan_define_event_graph": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int,
                                             C.POINTER(C.c_int32), C.c_int]),
    "sf_plan_finalize": (C.c_int, [C.c_void_p]),
    "sf_plan_smem_bytes": (C.c_int, [C.c_void_p, C.c_int]),
    "sf_plan_run_stage": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Event), C.c_void_p, C.c_void_p]),
    "sf_plan_run_events": (C.c_int, [C.c_void_p, C.POINTER(Event), C.c_int, C.c_void_p, C.c_void_p]),
    "sf_plan_last_launches": (C.c_int, [C.c_void_p]),
    "sf_plan_se_reduce": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Event), C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "sf_plan_se_apply": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Event), C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "sf_plan_se_reduce_totals": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Event), C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "sf_plan_se_totals_ptr": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "sf_plan_se_finish": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Event), C.c_void_p, C.c_float, C.c_void_p]),
    "sf_halo_copy": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                               C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "sf_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_ubyte)]),
    "sf_peer_open": (C.c_int, [C.POINTER(C.c_ubyte), C.POINTER(C.c_void_p)]),
    "sf_peer_close": (C.c_int, [C.c_void_p]),
    "sf_peer_free": (C.c_int, [C.c_void_p]),
    "sf_halo_push": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                               C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sf_halo_pull": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                               C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sf_peer_allreduce_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "sf_pack_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sf_unpack_nhwc_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sf_normal_policy": (C.c_int, [C.c_longlong, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sf_normal_fill_slots": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_ulonglong, C.c_ulonglong, C.c_int, C.c_int, C.c_void_p]),
    "sf_normal_fill_slot_list": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_ulonglong, C.c_ulonglong, C.c_int, C.c_int, C.c_void_p]),
    "sf_maxpool2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sf_upsample2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sf_cast_nhwc_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sf_space_to_depth2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sf_bilinear_up2_add": (C.c_int, [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sf_head_1x1": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                              C.c_void_p]),
    "sf_dwconv7_ln": (C.c_int, [C.c_void_p] * 8 + [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sf_aspp_pool_bias": (C.c_int, [C.c_void_p] * 8 + [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sf_pack_cell_weights": (C.c_int, [C.POINTER(Tensor), C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "sf_pack_pmodel_weights": (C.c_int, [C.POINTER(Tensor), C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "sf_packed_count": (C.c_int, [C.c_void_p]),
    "sf_packed_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(StageDesc)]),
    "sf_packed_free": (C.c_int, [C.c_void_p]),
    "sf_ode_query_workspace": (C.c_int, [C.POINTER(Geometry), C.POINTER(OdeOptions), C.POINTER(C.c_size_t)]),
    "sf_ode_create": (C.c_int, [C.POINTER(Geometry), C.POINTER(OdeOptions), C.POINTER(Tensor), C.c_int, C.c_char_p, C.c_void_p, C.c_size_t,
                                C.POINTER(C.c_void_p)]),
    "sf_ode_destroy": (C.c_int, [C.c_void_p]),
    "sf_ode_plan": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "sf_ode_tensor": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "sf_ode_set_observations": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "sf_ode_reset_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sf_ode_event": (C.c_int, [C.c_void_p, C.POINTER(Event), C.c_void_p, C.c_void_p]),
    "sf_ode_rollout": (C.c_int, [C.c_void_p, C.POINTER(Event), C.c_int, C.c_void_p, C.c_void_p]),
    "sf_ode_read_path": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "sf_merge_observations": (C.c_int, [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "sf_rollout_plan_create": (C.c_int, [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "sf_rollout_plan_info": (C.c_int, [C.c_void_p, C.POINTER(RolloutInfo)]),
    "sf_rollout_plan_events": (C.POINTER(Event), [C.c_void_p]),
    "sf_rollout_plan_table": (C.POINTER(C.c_int32), [C.c_void_p]),
    "sf_rollout_plan_out_slots": (C.POINTER(C.c_int32), [C.c_void_p]),
    "sf_rollout_plan_free": (C.c_int, [C.c_void_p]),
    "sf_diag_tma_dump": (C.c_int, [C.c_void_p] + [C.c_int] * 9 + [C.c_void_p, C.c_void_p]),
    "sf_diag_umma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "sf_diag_umma_ts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "sf_diag_umma_shift": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


class SfError(RuntimeError):
    pass


def load():
    """Loads the shared library once and types every export; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SfError(f"{LIB_PATH} not found: the CUDA library is not built (python -m streamingflow_b200.build). "
                      "streamingflow_b200 has no CPU / PyTorch fallback for the ODE path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)       # AttributeError if a declared symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.sf_abi_version() != SF_ABI_VERSION:
        raise SfError("libsf_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc < 0:
        msg = load().sf_last_error()
        raise SfError(f"{what}: {msg.decode() if msg else rc}")
    return rc
