"""ctypes binding of include/midi_emotion_b200.h (the C-ABI shared library).

There is no CPU fallback: if the library cannot be built/loaded, importing the compute path
raises.  Struct mirrors are checked against the library's own sizeof() at load time.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

ME_F32, ME_BF16 = 0, 1
COND_MODES = {"none": 0, "discrete_token": 1, "continuous_token": 2, "continuous_concat": 3}
ATTN_SIMT, ATTN_TENSOR = 0, 1
ATTN_NONCAUSAL, ATTN_REF_ROUNDING = 1, 2
EPI_BIAS, EPI_RELU, EPI_ADD_F32, EPI_RELU_MASK = 1, 2, 4, 8

_vp, _i32, _i64, _f32, _u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64


class AttnArgs(C.Structure):
    _fields_ = [
        ("dtype", _i32), ("impl", _i32),
        ("B", _i32), ("H", _i32), ("Lq", _i32), ("Lk", _i32), ("dh", _i32), ("max_seq", _i32),
        ("q_pos0", _i32), ("flags", _i32),
        ("q", _vp), ("k", _vp), ("v", _vp), ("E", _vp),
        ("q_sb", _i64), ("q_sh", _i64), ("q_si", _i64),
        ("k_sb", _i64), ("k_sh", _i64), ("k_sj", _i64),
        ("v_sb", _i64), ("v_sh", _i64), ("v_sj", _i64),
        ("keypad", _vp), ("keypad_ld", _i64),
        ("out", _vp), ("o_sb", _i64), ("o_si", _i64),
        ("lse", _vp), ("pos_dev", _vp), ("stream", _vp), ("p_tiles", _vp), ("m_tiles", _vp),
    ]


class AttnBwdArgs(C.Structure):
    _fields_ = [("f", AttnArgs), ("dout", _vp), ("dq", _vp), ("dk", _vp), ("dv", _vp), ("dE", _vp), ("dsum", _vp), ("dq_acc", _vp)]


_LAYER_PTRS = ["x_f32", "x_T", "keypad", "Wqkv", "bqkv", "E", "Wo", "bo", "ln1_w", "ln1_b", "W1", "b1", "W2", "b2",
               "ln2_w", "ln2_b", "qkv", "attn_o", "lse", "proj", "z1", "mean1", "rstd1", "out1_f32", "out1_T", "h",
               "z2", "mean2", "rstd2", "out2_f32", "out2_T", "stream", "attn_p", "attn_m", "xin_mean", "xin_rstd",
               "xin_gamma", "xin_beta"]


class ConvertDesc(C.Structure):
    _fields_ = [("src", _vp), ("dst", _vp), ("rows", _i32), ("cols", _i32), ("ld_src", _i32), ("ld_dst", _i32),
                ("src_dtype", _i32), ("dst_dtype", _i32)]


class SampleArgs(C.Structure):
    _fields_ = [("B", _i32), ("V", _i32), ("ld_logits", _i32), ("logits_dtype", _i32), ("logits", _vp),
                ("exclude", _vp), ("is_timeshift", _vp), ("prev_tokens", _vp), ("temp_note", _f32),
                ("temp_rest", _f32), ("penalty_coeff", _f32), ("top_k", _i32), ("top_p", _f32), ("_pad", _i32),
                ("repeat_counts", _vp), ("uniforms", _vp), ("out_tokens", _vp), ("out_num_choices", _vp),
                ("out_probs", _vp), ("stream", _vp)]


class TokenPipelineArgs(C.Structure):
    _fields_ = [("B", _i32), ("max_events", _i32), ("input_len", _i32), ("target_left_pad", _i32), ("pad_token", _i32),
                ("n_event_types", _i32), ("n_values", _i32), ("min_pitch", _i32), ("max_pitch", _i32), ("_pad", _i32),
                ("events", _vp), ("n_events", _vp), ("n_transpose", _vp), ("start", _vp), ("n_prefix", _vp),
                ("prefix", _vp), ("transposable", _vp), ("lut", _vp), ("input", _vp), ("target", _vp), ("status", _vp),
                ("stream", _vp)]


TP_MAX_PREFIX = 4


class AdamTensor(C.Structure):
    _fields_ = [("param", _vp), ("grad", _vp), ("exp_avg", _vp), ("exp_avg_sq", _vp), ("numel", _i64)]


class LayerArgs(C.Structure):
    _fields_ = [
        ("dtype", _i32), ("attn_impl", _i32), ("training", _i32), ("attn_flags", _i32),
        ("B", _i32), ("Ls", _i32), ("d", _i32), ("H", _i32), ("d_inner", _i32), ("max_seq", _i32),
        ("dropout_p", _f32), ("ln_eps", _f32), ("seed", _u64),
    ] + [(n, _vp) for n in _LAYER_PTRS]


_LAYER_BWD_PTRS = ["d_out", "d_x", "dWqkv", "dbqkv", "dE", "dWo", "dbo", "dln1_w", "dln1_b", "dW1", "db1", "dW2",
                   "db2", "dln2_w", "dln2_b", "g_a", "g_b", "g_T", "g_h", "g_qkv", "g_o", "dsum", "attn_ws",
                   "d_out_T", "d_x_T"]


class LayerBwdArgs(C.Structure):
    _fields_ = [("f", LayerArgs)] + [(n, _vp) for n in _LAYER_BWD_PTRS]


class DecodeLayerArgs(C.Structure):
    _fields_ = [("f", LayerArgs), ("k_cache", _vp), ("v_cache", _vp), ("t_dev", _vp), ("T_max", _i32),
                ("_pad", _i32)]


_PROTOS = {
    "me_last_error": (C.c_char_p, []),
    "me_version": (C.c_int, []),
    "me_launch_count": (C.c_ulonglong, []),
    "me_device_is_sm100": (C.c_int, []),
    "me_profile_enable": (C.c_int, [C.c_int]),
    "me_profile_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "me_profile_collect_class": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "me_debug_trace_set": (C.c_int, [_vp]),
    "me_sample_step": (C.c_int, [_vp]),
    "me_cross_entropy": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int64, _vp, C.c_int, _vp, _vp]),
    "me_sizeof_sample_args": (C.c_int, []),
    "me_sizeof_attn_args": (C.c_int, []),
    "me_sizeof_attn_bwd_args": (C.c_int, []),
    "me_sizeof_layer_args": (C.c_int, []),
    "me_sizeof_layer_bwd_args": (C.c_int, []),
    "me_sizeof_decode_layer_args": (C.c_int, []),
    "me_embed_forward": (C.c_int, [_vp] * 8 + [C.c_int] * 7 + [_f32, _u64, C.c_int, _vp, _vp, _vp, _vp]),
    "me_embed_backward": (C.c_int, [_vp] * 3 + [C.c_int] * 7 + [_f32, _u64] + [_vp] * 6),
    "me_embed_backward_split": (C.c_int, [_vp] * 4 + [C.c_int] * 7 + [_f32, _u64] + [_vp] * 6),
    "me_embed_decode": (C.c_int, [_vp] * 6 + [C.c_int] * 6 + [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp]),
    "me_gemm_bf16": (C.c_int, [_vp] * 3 + [C.c_int] * 10 + [_vp, _vp, _vp, C.c_int, _vp]),
    "me_gemm_bf16_ex": (C.c_int, [_vp] * 3 + [C.c_int] * 10 + [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "me_gemm_bf16_reference": (C.c_int, [_vp] * 3 + [C.c_int] * 8 + [_vp]),
    "me_gemm_f32": (C.c_int, [_vp] * 3 + [C.c_int] * 9 + [_vp, _vp, _vp, C.c_int, _vp]),
    "me_add_layernorm_forward": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _f32, C.c_int, C.c_int, _f32, _u64] + [_vp] * 6),
    "me_add_layernorm_backward": (C.c_int, [_vp] * 6 + [C.c_int, C.c_int, _f32, _u64, C.c_int] + [_vp] * 5),
    "me_colsum": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "me_colsum_ws": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int64, _vp]),
    "me_convert_2d": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "me_convert_batched": (C.c_int, [_vp, C.c_int, _vp]),
    "me_head_cross_entropy": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _i64, _vp, C.c_int,
                                        _vp, _vp]),
    "me_token_pipeline": (C.c_int, [C.POINTER(TokenPipelineArgs)]),
    "me_sizeof_token_pipeline_args": (C.c_int, []),
    "me_pooled_head_forward": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "me_pooled_head_backward": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp,
                                          _vp, _vp]),
    "me_attention_forward": (C.c_int, [C.POINTER(AttnArgs)]),
    "me_attention_backward": (C.c_int, [C.POINTER(AttnBwdArgs)]),
    "me_attention_backward_workspace_floats": (C.c_int64, [C.c_int] * 5),
    "me_attention_saved_tiles": (C.c_int64, [C.c_int, C.c_int]),
    "me_layer_forward": (C.c_int, [C.POINTER(LayerArgs)]),
    "me_layer_backward": (C.c_int, [C.POINTER(LayerBwdArgs)]),
    "me_decode_layer_step": (C.c_int, [C.POINTER(DecodeLayerArgs)]),
    "me_kv_cache_write": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp]),
    "me_cross_entropy_forward_backward": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _f32,
                                                    _vp, _vp, _vp, _vp]),
    "me_sample_step": (C.c_int, None),
    "me_grad_sqnorm_chunks": (C.c_int64, [C.POINTER(AdamTensor), C.c_int]),
    "me_grad_sqnorm_partials": (C.c_int, [C.POINTER(AdamTensor), C.c_int, C.c_double, _vp, C.c_int64, _vp]),
    "me_adam_prepare": (C.c_int, [_vp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, _vp, _vp, _vp]),
    "me_adam_update": (C.c_int, [C.POINTER(AdamTensor), C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                 C.c_double, _vp, _vp]),
}

# symbols every build must export (checked by the CPU test-suite against include/*.h)
REQUIRED_SYMBOLS = [k for k in _PROTOS if k not in ("me_cross_entropy_forward_backward", "me_sample_step")]

_lib = None
_lock = threading.Lock()


def library_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load (building first if necessary) the shared library.  Raises on failure: no fallback."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB
        if build_if_missing:
            try:
                path = _build.build()
            except Exception as e:  # stale-but-present library is still usable (GPU box has no reason to rebuild)
                if not os.path.exists(_build.LIB):
                    raise RuntimeError(f"midi_emotion_b200: CUDA library missing and cannot be built: {e}") from e
                path = _build.LIB
        if not os.path.exists(path):
            raise RuntimeError(f"midi_emotion_b200: {path} not found; run `python -m midi_emotion_b200.build`")
        lib = C.CDLL(path)
        for name, (res, args) in _PROTOS.items():
            if not hasattr(lib, name):
                if name in REQUIRED_SYMBOLS:
                    raise RuntimeError(f"midi_emotion_b200: library does not export {name}")
                continue
            fn = getattr(lib, name)
            fn.restype = res
            if args is not None:
                fn.argtypes = args
        for cname, st in (("me_sizeof_attn_args", AttnArgs), ("me_sizeof_attn_bwd_args", AttnBwdArgs),
                          ("me_sizeof_layer_args", LayerArgs), ("me_sizeof_layer_bwd_args", LayerBwdArgs),
                          ("me_sizeof_decode_layer_args", DecodeLayerArgs), ("me_sizeof_sample_args", SampleArgs),
                          ("me_sizeof_token_pipeline_args", TokenPipelineArgs)):
            if getattr(lib, cname)() != C.sizeof(st):
                raise RuntimeError(f"midi_emotion_b200: struct mirror {st.__name__} does not match the library")
        _lib = lib
        return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = _lib.me_last_error().decode("utf-8", "replace") if _lib is not None else "library not loaded"
        raise RuntimeError(f"midi_emotion_b200 {what}: {msg}")


def call(name: str, *args):
    """Invoke a C-ABI entry point, raising RuntimeError with me_last_error() on failure."""
    lib = _lib if _lib is not None else load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        check(rc, name)


def launch_count() -> int:
    """Kernels launched through the library so far in this process."""
    return int(load().me_launch_count())


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()
