"""ctypes binding of libmmdiff.so (C-ABI declared in include/mmdiff.h).

There is deliberately no fallback: if the shared library is missing or a call
fails, an exception is raised (the product path is the CUDA path or nothing).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMD_LIB", os.path.join(_HERE, "libmmdiff.so"))   # MMD_LIB: debug builds

MMD_MAX_LEVELS = 8


class MmdError(RuntimeError):
    pass


class MmdConfig(C.Structure):
    _fields_ = [
        ("video_f", C.c_int), ("video_c", C.c_int), ("video_h", C.c_int), ("video_w", C.c_int),
        ("audio_c", C.c_int), ("audio_l", C.c_int),
        ("model_channels", C.c_int),
        ("video_out_channels", C.c_int), ("audio_out_channels", C.c_int),
        ("num_res_blocks", C.c_int),
        ("n_levels", C.c_int),
        ("channel_mult", C.c_int * MMD_MAX_LEVELS),
        ("num_heads", C.c_int),
        ("num_head_channels", C.c_int),
        ("n_cross", C.c_int),
        ("cross_attention_resolutions", C.c_int * MMD_MAX_LEVELS),
        ("cross_attention_windows", C.c_int * MMD_MAX_LEVELS),
        ("cross_attention_shift", C.c_int),
        ("n_video_attn", C.c_int),
        ("video_attention_resolutions", C.c_int * MMD_MAX_LEVELS),
        ("n_audio_attn", C.c_int),
        ("audio_attention_resolutions", C.c_int * MMD_MAX_LEVELS),
        ("max_batch", C.c_int),
    ]


class MmdConvDesc(C.Structure):
    _fields_ = [
        ("rank", C.c_int),
        ("dims", C.c_int64 * 4),
        ("box", C.c_int * 4),
        ("n_src", C.c_int),
        ("src", C.c_void_p * 4),
        ("src_channels", C.c_int * 4),
        ("n_taps", C.c_int),
        ("taps", (C.c_int * 3) * 27),
        ("weight", C.c_void_p),
        ("bias", C.c_void_p),
        ("n", C.c_int),
        ("out", C.c_void_p),
        ("out_f32", C.c_void_p),
        ("ostride", C.c_int64 * 4),
        ("ostride_c", C.c_int64),
        ("gn_sums", C.c_void_p),
        ("gn_rows", C.c_int64),
    ]


class MmdAttnDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_ld", C.c_int), ("q_col0", C.c_int), ("q_rows", C.c_int64),
        ("k", C.c_void_p), ("k_ld", C.c_int), ("k_col0", C.c_int), ("k_rows", C.c_int64),
        ("v", C.c_void_p), ("v_ld", C.c_int), ("v_col0", C.c_int),
        ("out", C.c_void_p), ("out_ld", C.c_int),
        ("batch", C.c_int), ("heads", C.c_int), ("head_dim", C.c_int),
        ("n_blocks", C.c_int), ("q_blk", C.c_int), ("k_blk", C.c_int), ("win", C.c_int), ("shift", C.c_int),
    ]


# (name, restype, argtypes) — must list every symbol include/mmdiff.h declares.
_PROTOS = [
    ("mmd_last_error", C.c_char_p, []),
    ("mmd_version", C.c_char_p, []),
    ("mmd_model_create", C.c_int, [C.POINTER(MmdConfig), C.POINTER(C.c_void_p)]),
    ("mmd_model_destroy", C.c_int, [C.c_void_p]),
    ("mmd_model_num_params", C.c_int, [C.c_void_p]),
    ("mmd_model_param_info", C.c_int,
     [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    ("mmd_model_set_param", C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("mmd_model_num_shifts", C.c_int, [C.c_void_p]),
    ("mmd_model_shift_bound", C.c_int, [C.c_void_p, C.c_int]),
    ("mmd_model_workspace_bytes", C.c_size_t, [C.c_void_p, C.c_int]),
    ("mmd_model_num_launches", C.c_int, [C.c_void_p, C.c_int]),
    ("mmd_model_forward", C.c_int,
     [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p,
      C.c_void_p]),
    ("mmd_model_forward_train", C.c_int,
     [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p,
      C.c_void_p]),
    ("mmd_model_backward", C.c_int,
     [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("mmd_model_train_workspace_bytes", C.c_size_t, [C.c_void_p, C.c_int]),
    ("mmd_model_param_offset", C.c_int64, [C.c_void_p, C.c_int]),
    ("mmd_model_param_floats", C.c_int64, [C.c_void_p]),
    ("mmd_model_num_backward_launches", C.c_int, [C.c_void_p, C.c_int]),
    ("mmd_model_profile_backward", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_void_p]),
    ("mmd_model_backward_step_kind", C.c_char_p, [C.c_void_p, C.c_int, C.c_int]),
    ("mmd_model_profile", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_void_p]),
    ("mmd_model_step_info", C.c_int,
     [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_double),
      C.POINTER(C.c_int)]),
    ("mmd_p_sample_tail", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
      C.c_void_p]),
    ("mmd_q_sample", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    ("mmd_sample_epilogue", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p]),
    ("mmd_lincomb", C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_float), C.c_int64, C.c_void_p, C.c_void_p]),
    ("mmd_dpm_threshold", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_void_p]),
    ("mmd_dpm_error_sq", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_float, C.c_void_p,
                                   C.c_void_p]),
    ("mmd_op_group_norm", C.c_int,
     [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
      C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    ("mmd_op_group_norm_temporal", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("mmd_op_resample", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("mmd_op_conv", C.c_int, [C.POINTER(MmdConvDesc), C.c_void_p]),
    ("mmd_op_conv_timed", C.c_int, [C.POINTER(MmdConvDesc), C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    ("mmd_model_set_params_flat", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("mmd_model_set_dropout", C.c_int, [C.c_void_p, C.c_float, C.c_uint64]),
    ("mmd_model_train_generation", C.c_int64, [C.c_void_p, C.c_int]),
    ("mmd_model_num_dropout_sites", C.c_int, [C.c_void_p, C.c_int]),
    ("mmd_model_dropout_site", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("mmd_model_dropout_mask", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    ("mmd_op_conv_gn", C.c_int, [C.POINTER(MmdConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p]),
    ("mmd_op_attention", C.c_int, [C.POINTER(MmdAttnDesc), C.c_void_p]),
    ("mmd_op_temporal_attention", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("mmd_op_conv_wgrad", C.c_int, [C.POINTER(MmdConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("mmd_op_conv_dgrad", C.c_int, [C.POINTER(MmdConvDesc), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    ("mmd_op_group_norm_bwd", C.c_int,
     [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
      C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("mmd_op_group_norm_temporal_bwd", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("mmd_op_resample_bwd", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("mmd_op_temporal_attention_bwd", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("mmd_op_attention_fwd_bwd", C.c_int,
     [C.POINTER(MmdAttnDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
      C.c_void_p]),
    ("mmd_op_head_bwd", C.c_int, [C.POINTER(MmdConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
]

EXPORTED_SYMBOLS = [p[0] for p in _PROTOS]

_lib = None


def load() -> C.CDLL:
    """Load libmmdiff.so; raises MmdError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MmdError(
            f"{LIB_PATH} is missing: build it with `python -m mm_diffusion_b200.build` "
            "(__graft_entry__.build()). There is no CPU/PyTorch fallback for the denoising path.")
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in _PROTOS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported (no partial library)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        msg = load().mmd_last_error()
        raise MmdError(f"libmmdiff error {code}: {msg.decode() if msg else '?'}")


def ptr(t) -> int:
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
