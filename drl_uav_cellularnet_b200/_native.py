"""ctypes binding of libuavenv.so (include/uavenv.h).  There is no fallback: if the library is missing the
import fails loudly -- build it with ``python -m drl_uav_cellularnet_b200.build``."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# UAVENV_SO selects another build of the same library (tuning variants made by build.build(defines=..., out=...))
SO = os.environ.get("UAVENV_SO") or os.path.join(PKG, "libuavenv.so")

MAX_BS = 32
MAX_GROUPS = 32

MOB_GROUP, MOB_TRACE = 0, 1
FADE_PHILOX, FADE_INJECTED, FADE_NONE = 0, 1, 2
PREC_FP32_FAST, PREC_FP64_PARITY, PREC_FP32_GUARDED = 0, 1, 2
OBS_NONE, OBS_F32, OBS_F32_INCREMENTAL = 0, 1, 3
OK, EINVAL, ECUDA, ETRACE, ENOMEM, EACTION = 0, -1, -2, -3, -4, -5
ERR_ACTION, ERR_TRACE, ERR_CLAMP = 1, 2, 4


class Cfg(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("n_bs", C.c_int32), ("n_ue", C.c_int32), ("grid_n", C.c_int32),
        ("mobility", C.c_int32), ("fading", C.c_int32), ("precision", C.c_int32), ("obs_mode", C.c_int32),
        ("seed", C.c_uint64), ("env_offset", C.c_int64),
        ("device", C.c_int32), ("max_step", C.c_int32), ("n_act", C.c_int32), ("bs_step", C.c_int32),
        ("min_bs_dist", C.c_int32), ("warmup_ticks", C.c_int32), ("n_groups", C.c_int32),
        ("group_sizes", C.c_int32 * MAX_GROUPS), ("has_init_bs", C.c_int32), ("init_bs_xy", C.c_int32 * (MAX_BS * 2)),
        ("aggregating0", C.c_int32), ("deaggregating0", C.c_int32),
        ("deaggregating_len", C.c_int32), ("aggregating_len", C.c_int32),
        ("grid_width", C.c_double), ("p_bs_dbm", C.c_double), ("noise_dbm", C.c_double),
        ("pl_a", C.c_double), ("pl_b", C.c_double), ("pl_dis", C.c_double),
        ("ant_gain", C.c_double), ("eq_loss", C.c_double),
        ("shadow_mean", C.c_double), ("shadow_sd", C.c_double),
        ("ho_thresh_db", C.c_double), ("out_thresh_db", C.c_double),
        ("v_min", C.c_double), ("v_max", C.c_double), ("aggregation", C.c_double),
        ("guard_db", C.c_double),
    ]


class In(C.Structure):
    _fields_ = [("action", C.c_void_p), ("digits", C.c_void_p), ("fading", C.c_void_p),
                ("mob_uniforms", C.c_void_p), ("env_mask", C.c_void_p)]


class Out(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("reward", C.c_void_p), ("mean_sinr", C.c_void_p), ("n_out", C.c_void_p),
                ("n_ho", C.c_void_p), ("n_blocked", C.c_void_p), ("done", C.c_void_p), ("step_n", C.c_void_p),
                ("serving", C.c_void_p), ("serving_sinr", C.c_void_p), ("sinr_all", C.c_void_p),
                ("fading_used", C.c_void_p), ("ue_xy", C.c_void_p), ("bs_xy", C.c_void_p), ("bs_digits", C.c_void_p),
                ("obs_idx", C.c_void_p)]


class GemmDesc(C.Structure):
    """uavnet_gemm_desc (include/uavnet.h)"""
    _fields_ = [("A", C.c_void_p), ("lda", C.c_int64), ("a_trans", C.c_int32),
                ("B", C.c_void_p), ("ldb", C.c_int64), ("b_trans", C.c_int32),
                ("D", C.c_void_p), ("ldd", C.c_int64),
                ("M", C.c_int64), ("N", C.c_int32), ("K", C.c_int64),
                ("bias", C.c_void_p), ("relu6", C.c_int32),
                ("mask_src", C.c_void_p), ("ld_mask", C.c_int64),
                ("accumulate", C.c_int32), ("split_k", C.c_int32),
                ("colsum", C.c_void_p), ("out_colsum", C.c_void_p),
                ("dot_w", C.c_void_p), ("dot_b", C.c_void_p), ("dot_out", C.c_void_p),
                ("precision", C.c_int32)]


class PushPart(C.Structure):
    """uavnet_push_part (include/uavnet.h): a column range of a row-major matrix inside the flat buffers, in float32 elements"""
    _fields_ = [("offset", C.c_int64), ("rows", C.c_int64), ("row_width", C.c_int32), ("col0", C.c_int32), ("n_cols", C.c_int32)]


GEMM_TF32, GEMM_3XTF32 = 0, 1

# every symbol include/uavenv.h and include/uavnet.h declare
SYMBOLS = [
    "uavenv_cfg_default", "uavenv_create", "uavenv_destroy", "uavenv_set_trace", "uavenv_ctor_pass",
    "uavenv_reset", "uavenv_step", "uavenv_step_host", "uavenv_step_host_state", "uavenv_coverage_map", "uavenv_state_bytes", "uavenv_state_field",
    "uavenv_get_state", "uavenv_set_state", "uavenv_check", "uavenv_guard_hits", "uavenv_get_cfg", "uavenv_last_error",
    "uavnet_sparse_fwd", "uavnet_sparse_bwd", "uavnet_sparse_bwd_gather", "uavnet_sparse_bwd_gather_workspace", "uavnet_sparse_bwd_gather_prepare", "uavnet_sparse_bwd_gather_apply", "uavnet_sparse_bwd_gather_apply_cols", "uavnet_rmsprop", "uavnet_actor_head_bwd", "uavnet_softmax_sample", "uavnet_p2p_alloc", "uavnet_p2p_open", "uavnet_p2p_close", "uavnet_p2p_free",
    "uavnet_p2p_rmsprop", "uavnet_p2p_push", "uavnet_p2p_push_status", "uavnet_p2p_push_part", "uavnet_gemm", "uavnet_gemm_check", "uavnet_nstep_targets", "uavnet_rank1_mask", "uavnet_rollout_record", "uavnet_critic_td", "uavnet_mean_rows",
    "uavenv_launch_count", "uavenv_version", "uavenv_launch_plan",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(SO):
        raise ImportError(
            "drl_uav_cellularnet_b200: %s is missing. There is no CPU / eager fallback; build the CUDA library "
            "with `python -m drl_uav_cellularnet_b200.build`." % SO)
    L = C.CDLL(SO)
    P, vp = C.POINTER, C.c_void_p
    L.uavenv_cfg_default.argtypes = [P(Cfg), C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    L.uavenv_create.argtypes = [P(Cfg), P(vp)]
    L.uavenv_destroy.argtypes = [vp]
    L.uavenv_destroy.restype = None
    L.uavenv_set_trace.argtypes = [vp, vp, C.c_int64, C.c_int32]
    for f in (L.uavenv_ctor_pass, L.uavenv_reset, L.uavenv_step):
        f.argtypes = [vp, P(In), P(Out), vp]
    L.uavenv_step_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.uavenv_step_host_state.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.uavenv_coverage_map.argtypes = [vp, vp, vp, vp, vp]
    L.uavenv_state_bytes.argtypes = [vp]
    L.uavenv_state_bytes.restype = C.c_int64
    L.uavenv_state_field.argtypes = [vp, C.c_int32, P(C.c_int64), P(C.c_int64)]
    L.uavenv_get_state.argtypes = [vp, vp, C.c_int64]
    L.uavenv_set_state.argtypes = [vp, vp, C.c_int64]
    L.uavenv_check.argtypes = [vp, P(C.c_uint32), vp]
    L.uavenv_guard_hits.argtypes = [vp, P(C.c_int64), vp]
    L.uavenv_get_cfg.argtypes = [vp]
    L.uavenv_get_cfg.restype = P(Cfg)
    L.uavenv_last_error.argtypes = [vp]
    L.uavenv_last_error.restype = C.c_char_p
    L.uavenv_launch_count.argtypes = [vp]
    L.uavenv_launch_count.restype = C.c_int64
    L.uavenv_version.restype = C.c_char_p
    L.uavenv_launch_plan.argtypes = [vp, P(C.c_int32), P(C.c_int32), P(C.c_int32), P(C.c_int32)]
    L.uavnet_sparse_fwd.argtypes = [vp, C.c_int64, C.c_int32, C.c_int64, vp, vp, C.c_int32, vp, C.c_int32, vp]
    L.uavnet_sparse_bwd.argtypes = [vp, C.c_int64, C.c_int32, C.c_int64, vp, C.c_int32, vp, vp]
    L.uavnet_sparse_bwd_gather.argtypes = [vp, C.c_int64, C.c_int32, C.c_int64, vp, C.c_int32, vp, vp, C.c_int32, vp]
    L.uavnet_sparse_bwd_gather_prepare.argtypes = [vp, C.c_int64, C.c_int32, C.c_int64, vp, vp]
    L.uavnet_sparse_bwd_gather_apply.argtypes = [C.c_int64, C.c_int32, C.c_int64, vp, C.c_int32, vp, vp, C.c_int32, vp]
    L.uavnet_sparse_bwd_gather_apply_cols.argtypes = [C.c_int64, C.c_int32, C.c_int64, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp]
    L.uavnet_sparse_bwd_gather_workspace.argtypes = [C.c_int64, C.c_int32, C.c_int64]
    L.uavnet_sparse_bwd_gather_workspace.restype = C.c_int64
    L.uavnet_softmax_sample.argtypes = [vp, C.c_int64, C.c_int32, C.c_uint64, C.c_uint32, vp, C.c_uint32, vp, vp, vp]
    L.uavnet_actor_head_bwd.argtypes = [vp, vp, vp, C.c_int64, C.c_int32, C.c_float, vp, C.c_int64, vp, vp]
    L.uavnet_p2p_alloc.argtypes = [C.c_int64, P(vp), vp]
    L.uavnet_p2p_open.argtypes = [vp, P(vp)]
    L.uavnet_p2p_close.argtypes = [vp]
    L.uavnet_p2p_free.argtypes = [vp]
    L.uavnet_p2p_rmsprop.argtypes = [P(vp), P(vp), vp, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, vp]
    L.uavnet_p2p_push.argtypes = [P(vp), P(vp), P(vp), vp, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, vp]
    L.uavnet_p2p_push_status.argtypes = [vp, P(C.c_uint32), P(C.c_uint32)]
    L.uavnet_p2p_push_part.argtypes = [P(vp), P(vp), P(vp), vp, P(PushPart), P(PushPart), C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, vp]
    L.uavnet_rmsprop.argtypes = [vp, vp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, vp]
    L.uavnet_gemm.argtypes = [P(GemmDesc), vp]
    L.uavnet_gemm_check.argtypes = []
    L.uavnet_rank1_mask.argtypes = [vp, vp, vp, C.c_int64, C.c_int32, vp, vp]
    L.uavnet_rollout_record.argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp, vp]
    L.uavnet_critic_td.argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp]
    L.uavnet_mean_rows.argtypes = [vp, C.c_int64, vp, vp]
    L.uavnet_nstep_targets.argtypes = [vp, vp, vp, C.c_int32, C.c_int64, C.c_float, vp, vp]
    _lib = L
    return L


# ---- diagnostics library (include/uavenv_diag.h): not part of the product boundary ----
DIAG_SO = os.path.join(PKG, "libuavenv_diag.so")
DIAG_SYMBOLS = ["uavenv_diag_fill", "uavenv_diag_fill_ring", "uavenv_diag_fill_env"]
_diag = None


def diag_lib():
    global _diag
    if _diag is None:
        if not os.path.isfile(DIAG_SO):
            raise ImportError("%s is missing: build it with `python -m drl_uav_cellularnet_b200.build`" % DIAG_SO)
        L = C.CDLL(DIAG_SO)
        vp = C.c_void_p
        L.uavenv_diag_fill.argtypes = [vp, C.c_int64, C.c_int64, C.c_int32, vp]
        L.uavenv_diag_fill_ring.argtypes = [vp, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
        L.uavenv_diag_fill_env.argtypes = [vp, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
        _diag = L
    return _diag
