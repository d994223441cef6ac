"""ctypes binding of libscb200.so (the C ABI declared in include/speechcatcher_b200.h).

The CUDA path has no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libscb200.so"


class ScConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "d_model", "enc_heads", "enc_layers", "dec_heads", "dec_layers", "vocab", "ffn", "n_streams", "beam",
        "max_chunk", "max_frames", "use_bbd", "precision")] + [("ctc_weight", C.c_float)]


class ScPushStats(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_feature_frames", "n_encoder_blocks", "n_encoder_frames", "n_decode_steps", "n_kernel_launches")] + \
        [("reserved", C.c_int32 * 3)]


class ScStreamPlan(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "called", "n_feat", "n_sub", "n_blocks", "n_enc_out", "enc_len", "n_decode_blocks", "last_T")]


class ScSegmentParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("beam_size", "ideal_segment_len", "max_lookahead", "min_len", "step",
                                         "reserved")] + [("len_reward_weight", C.c_double), ("energy_weight", C.c_double)]


# every symbol include/speechcatcher_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _sz = C.c_void_p, C.c_int32, C.c_size_t
_pi32, _pf, _pd = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double)
SYMBOLS = {
    "sc_version": (C.c_char_p, []),
    "sc_last_error": (C.c_char_p, []),
    "sc_engine_workspace_bytes": (C.c_int, [C.POINTER(ScConfig), C.POINTER(_sz)]),
    "sc_engine_create": (C.c_int, [C.POINTER(ScConfig), _vp, _sz, C.POINTER(_vp)]),
    "sc_engine_destroy": (C.c_int, [_vp]),
    "sc_engine_set_weight": (C.c_int, [_vp, C.c_char_p, _vp, _sz]),
    "sc_engine_set_frontend": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "sc_engine_finalize": (C.c_int, [_vp]),
    "sc_engine_reset": (C.c_int, [_vp, _vp, _i32, _vp]),
    "sc_engine_push": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp, C.POINTER(ScPushStats)]),
    "sc_engine_push_features": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp, C.POINTER(ScPushStats)]),
    "sc_engine_read_beam": (C.c_int, [_vp, _i32, _i32, _pi32, _pi32, _pi32, _vp, _vp, _vp, _vp]),
    "sc_engine_read_all": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "sc_engine_token_capacity": (C.c_int, [_vp, _pi32]),
    "sc_engine_last_plan": (C.c_int, [_vp, _i32, C.POINTER(ScStreamPlan)]),
    "sc_engine_buffer": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp), C.POINTER(_sz)]),
    "sc_engine_trace_layout": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "sc_engine_set_trace": (C.c_int, [_vp, _vp, _sz, _i32]),
    "sc_engine_counter": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_int64)]),
    "sc_engine_set_option": (C.c_int, [_vp, C.c_char_p, _i32]),
    "sc_engine_profile_begin": (C.c_int, [_vp, _i32, _i32, _i32]),
    "sc_engine_profile_end": (C.c_int, [_vp, _i32, _pi32, _pd, _pd, C.POINTER(C.c_uint64)]),
    "sc_planner_create": (C.c_int, [_i32, C.POINTER(_vp)]),
    "sc_planner_destroy": (C.c_int, [_vp]),
    "sc_planner_reset": (C.c_int, [_vp, _i32]),
    "sc_planner_push": (C.c_int, [_vp, _i32, _i32, _i32, C.POINTER(ScStreamPlan)]),
    "sc_planner_push_features": (C.c_int, [_vp, _i32, _i32, _i32, C.POINTER(ScStreamPlan)]),
    "sc_segment_num_frames": (C.c_int, [C.c_int64, C.POINTER(C.c_int64)]),
    "sc_segment_filterbank_bins": (C.c_int, [_pd]),
    "sc_segment_energy": (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_int64, C.c_double, _vp]),
    "sc_segment_search": (C.c_int, [_pd, C.c_int64, C.POINTER(ScSegmentParams), C.POINTER(C.c_int64), _i32, _pi32]),
    "sc_frontend_workspace_bytes": (C.c_size_t, []),
    "sc_frontend_init": (C.c_int, [_vp, _vp, _vp]),
    "sc_frontend_fbank_mvn": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _pi32, _vp]),
    "sc_ctc_prefix_step": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "sc_layernorm_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "sc_linear_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "sc_linear_x3": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "sc_linear_x3_ln": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, _i32, _i32, _i32, _vp, _vp]),
    "sc_linear_x3_planes": (C.c_int, [_vp, C.c_int64, _i32, _vp, _vp, _vp, _vp, _vp, C.c_int64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "sc_layernorm_split": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, _i32, _i32, _vp]),
    "sc_linear_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "sc_linear_bf16_lnA": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "sc_ffn_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "sc_ffn_bf16_timeline": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "sc_linear_bf16_ln": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
}

_lib = None


def load():
    """Load libscb200.so (building is a separate, explicit step: speechcatcher_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension must be built first "
            "(python -m speechcatcher_b200.build); there is no CPU fallback")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # raises AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().sc_last_error().decode(errors="replace")
        raise RuntimeError(f"speechcatcher_b200 {what} failed (code {rc}): {msg}")
