"""ctypes binding of libtheia_b200.so (include/theia_b200.h).  Thin: no pybind, no torch types in
the ABI -- tensors cross as raw device pointers + the current CUDA stream handle.

The library is REQUIRED: there is no Python / PyTorch fallback for any op.  If it is missing or a
call fails this module raises."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# THEIA_B200_LIB: A/B experiments against an alternative in-tree build (tools only; the product path is the default)
LIB_PATH = os.environ.get("THEIA_B200_LIB") or os.path.join(_PKG, "libtheia_b200.so")

OP_K2D, OP_MN2D, OP_CONV_K, OP_CONV_MN = 0, 1, 2, 3
EPI_GELU, EPI_RELU, EPI_RESID, EPI_OUT_F32, EPI_ATOMIC = 1 << 1, 1 << 2, 1 << 3, 1 << 4, 1 << 5
EPI_MUL_AUX, EPI_MUL_RELUMASK, EPI_POSCLS, EPI_STATS, EPI_COLSUM = 1 << 6, 1 << 7, 1 << 8, 1 << 9, 1 << 10
EPI_GELU_FWD, EPI_QUICK_GELU, EPI_RESID_F32, EPI_AUX_U8 = 1 << 11, 1 << 12, 1 << 13, 1 << 14
MAX_TEACHERS = 8


class ConvGeom(C.Structure):
    _fields_ = [("C", C.c_int), ("H", C.c_int), ("W", C.c_int), ("B", C.c_int),
                ("stride_w", C.c_longlong), ("stride_h", C.c_longlong), ("stride_b", C.c_longlong),
                ("ntaps", C.c_int), ("dh", C.c_int * 9), ("dw", C.c_int * 9),
                ("tile_w", C.c_int), ("tile_h", C.c_int), ("out_h", C.c_int), ("out_w", C.c_int),
                ("out_img_rows", C.c_int), ("out_row_off", C.c_int), ("out_wpitch", C.c_int),
                ("sy", C.c_int), ("sx", C.c_int), ("py", C.c_int), ("px", C.c_int),
                ("in_stride", C.c_int), ("b_tap_rows", C.c_int), ("wtap", C.c_int * 9)]


class GemmDesc(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("a_mode", C.c_int), ("b_mode", C.c_int),
                ("A", C.c_void_p), ("lda", C.c_longlong), ("B", C.c_void_p), ("ldb", C.c_longlong),
                ("conv", ConvGeom), ("epi", C.c_int), ("out", C.c_void_p), ("ldo", C.c_longlong),
                ("out2", C.c_void_p), ("bias", C.c_void_p), ("aux", C.c_void_p), ("pos", C.c_void_p),
                ("cls", C.c_void_p), ("tokens", C.c_int), ("stats", C.c_void_p), ("rows_per_image", C.c_int),
                ("splits", C.c_int), ("batch_z", C.c_int), ("out_z_stride", C.c_longlong), ("bn", C.c_int), ("colsum", C.c_void_p),
                ("tok_p0", C.c_int), ("tok_p1", C.c_int)]


class ModelConfig(C.Structure):
    _fields_ = [("hidden", C.c_int), ("heads", C.c_int), ("layers", C.c_int), ("image", C.c_int),
                ("patch", C.c_int), ("max_batch", C.c_int), ("ln_eps", C.c_float), ("variant", C.c_int),
                ("num_reg_tokens", C.c_int), ("num_teachers", C.c_int),
                ("teacher_names", C.c_char_p * MAX_TEACHERS), ("teacher_c", C.c_int * MAX_TEACHERS),
                ("teacher_hw", C.c_int * MAX_TEACHERS)]


class VitLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ln1_w", "ln1_b", "w_qkv", "b_qkv", "w_o", "b_o", "ln2_w", "ln2_b",
                                          "w_fc1", "b_fc1", "w_fc2", "b_fc2")]


class VitDesc(C.Structure):
    _fields_ = [("hidden", C.c_int), ("heads", C.c_int), ("layers", C.c_int), ("mlp", C.c_int),
                ("tokens", C.c_int), ("patch_off", C.c_int), ("patch_tokens", C.c_int), ("patch_k", C.c_int),
                ("ln_eps", C.c_float), ("act", C.c_int),
                ("w_patch", C.c_void_p), ("b_patch", C.c_void_p), ("tok_table", C.c_void_p),
                ("pre_ln_w", C.c_void_p), ("pre_ln_b", C.c_void_p), ("final_ln_w", C.c_void_p), ("final_ln_b", C.c_void_p),
                ("final_ln_mode", C.c_int), ("layer", C.POINTER(VitLayer)), ("residual_f32", C.c_int)]


# every symbol include/theia_b200.h declares: name -> (restype, argtypes)
_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float
SYMBOLS = {
    "theia_last_error": (C.c_char_p, []),
    "theia_version": (_i, []),
    "theia_launch_count": (_ll, []),
    "theia_gemm": (_i, [C.POINTER(GemmDesc), _vp]),
    "theia_debug_set": (_i, [_i, _ll]),
    "theia_plan_wgrad": (_i, [_i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "theia_prof_enable": (_i, [_i]),
    "theia_prof_record": (_i, [_i, C.POINTER(C.c_double), C.POINTER(_i)]),
    "theia_prof_collect": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_ll)]),
    "theia_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "theia_layernorm_fwd_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "theia_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "theia_ln3d_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _f, _i, _i, _i, _vp]),
    "theia_ln3d_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _i, _i, _i, _i, _vp]),
    "theia_adamw_flat": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _f, _vp, _vp, _vp, _vp]),
    "theia_pack_cast": (_i, [_vp, _vp, _vp, _ll, _vp]),
    "theia_conv_pack": (_i, [_vp, _i, _i, _vp]),
    "theia_conv_unpack": (_i, [_vp, _i, _i, _vp, _vp]),
    "theia_perm_segments": (_i, [_vp, _i, _ll, _vp, _vp]),
    "theia_target_ingest": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "theia_chw_to_hwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "theia_hwc_to_chw": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "theia_loss_fwd": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "theia_loss_bwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _vp]),
    "theia_preprocess": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float), _i, _i, _vp]),
    "theia_preprocess_hw": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float), _i, _i, _vp]),
    "theia_preprocess_f32": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float), _i, _i, _vp]),
    "theia_preprocess_debug_u8": (_i, [_vp]),
    "theia_attention_tc_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "theia_attention_tc_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "theia_attention_fwd_hd80": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "theia_gather4": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _ll, _vp]),
    "theia_cast_bf16": (_i, [_vp, _vp, _ll, _vp]),
    "theia_transpose_cast_bf16": (_i, [_vp, _vp, _i, _i, _vp]),
    "theia_colsum_tokens": (_i, [_vp, _vp, _i, _i, _ll, _i, _i, _i, _vp]),
    "theia_colsum": (_i, [_vp, _vp, _i, _i, _ll, _i, _vp]),
    "theia_batchsum": (_i, [_vp, _vp, _i, _i, _vp]),
    "theia_model_create": (_i, [C.POINTER(ModelConfig), C.POINTER(_vp)]),
    "theia_model_destroy": (None, [_vp]),
    "theia_model_param_floats": (_ll, [_vp]),
    "theia_model_workspace_bytes": (_ll, [_vp]),
    "theia_model_num_params": (_i, [_vp]),
    "theia_model_param_info": (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(_ll), C.POINTER(_i), C.POINTER(_ll)]),
    "theia_model_debug_ptr": (_i, [_vp, C.c_char_p, _i, C.POINTER(_vp), C.POINTER(_ll), C.POINTER(_i)]),
    "theia_model_bind": (_i, [_vp, _vp, _vp, _vp]),
    "theia_model_pack": (_i, [_vp, _i, _vp]),
    "theia_model_pack_table": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "theia_model_set_grads": (_i, [_vp, _vp]),
    "theia_model_set_input_size": (_i, [_vp, _i, _i]),
    "theia_model_set_input_dtype": (_i, [_vp, _i]),
    "theia_model_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float), _i,
                                 C.POINTER(_vp), _vp, _vp]),
    "theia_model_backward": (_i, [_vp, C.POINTER(_vp), _vp]),
    "theia_vit_workspace_bytes": (_ll, [C.POINTER(VitDesc), _i]),
    "theia_vit_forward": (_i, [C.POINTER(VitDesc), _vp, _i, _vp, _vp, _vp, _vp]),
    "theia_patchify_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
}

_lib = None


class TheiaError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TheiaError(
                f"{LIB_PATH} is missing: build it with `python -m theia_b200._build` (or __graft_entry__.build()). "
                "theia_b200 has no CPU / PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().theia_last_error()
        raise TheiaError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
