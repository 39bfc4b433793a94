"""ctypes binding of libcsd_b200.so (the C ABI in include/csd_b200.h).

This is the only place the shared library is opened. There is NO fallback: if the library is
missing or an entry point fails, the caller gets an exception. The reference binds its native
code with torch.utils.cpp_extension.load at import (op/upfirdn2d.py:10-16, op/fused_act.py:11-17);
here the library is prebuilt in-tree by `conditional_score_diffusion_b200.build`.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# CSD_LIB_PATH: another build of the same library (A/B measurements of two kernels versions in one GPU call)
LIB_PATH = os.environ.get("CSD_LIB_PATH") or os.path.join(_HERE, "libcsd_b200.so")

CSD_MAX_SEGMENTS = 4

c_int = ctypes.c_int
c_int32 = ctypes.c_int32
c_int64 = ctypes.c_int64
c_float = ctypes.c_float
c_void_p = ctypes.c_void_p


class CsdError(RuntimeError):
    """A libcsd_b200 entry point returned a non-zero status."""


class ConvSegment(ctypes.Structure):
    _fields_ = [
        ("a", c_void_p),
        ("pitch", c_int32),
        ("c_off", c_int32),
        ("c_cnt", c_int32),
        ("taps", c_int32),
        ("norm", c_void_p),
        ("norm_silu", c_int32),
        ("reserved_", c_int32),
    ]


class ConvGemmDesc(ctypes.Structure):
    _fields_ = [
        ("batch", c_int32), ("h", c_int32), ("w", c_int32),
        ("in_h", c_int32), ("in_w", c_int32), ("stride", c_int32), ("pad", c_int32),
        ("tile_w", c_int32), ("tile_h", c_int32), ("tile_b", c_int32),
        ("mode", c_int32), ("mt", c_int32),
        ("nseg", c_int32),
        ("n", c_int32),
        ("n_store", c_int32),
        ("n_tile", c_int32),
        ("seg", ConvSegment * CSD_MAX_SEGMENTS),
        ("wt", c_void_p),
        ("wt_rows", c_int32),
        ("k_total", c_int32),
        ("wt_pitch", c_int32),
        ("wt_k_off", c_int32),
        ("k_valid", c_int32),
        ("wt_batch_stride", c_int64),
        ("z_batches", c_int32),
        ("a_batch_step", c_int32),
        ("out", c_void_p),
        ("out_pitch", c_int32),
        ("out_f32", c_int32),
        ("out_z_stride", c_int64),
        ("bias", c_void_p),
        ("bias_per_row", c_int32),
        ("temb", c_void_p),
        ("temb_pitch", c_int32),
        ("res", c_void_p),
        ("res_pitch", c_int32),
        ("res_z_stride", c_int64),
        ("scale", c_float),
        ("stat_partials", c_void_p),
        ("dtype", c_int32),
        ("out_round_tf32", c_int32),
        ("k_splits", c_int32),
        ("splitk_ws", c_void_p),
    ]


class PackJob(ctypes.Structure):
    _fields_ = [("src", c_void_p), ("src2", c_void_p), ("dst", c_void_p), ("s_row", c_int64), ("s_col", c_int64),
                ("s_tap", c_int64), ("rows", c_int32), ("cols", c_int32), ("taps", c_int32), ("k_pad", c_int32),
                ("dst_pitch", c_int32), ("flip", c_int32), ("kind", c_int32), ("scale", c_float)]


class PixMajorGeom(ctypes.Structure):
    _fields_ = [("q", c_int64), ("kp", c_int64), ("row_pitch", c_int64), ("wp", c_int32), ("ips", c_int32),
                ("splits", c_int32), ("margin", c_int32)]


# name -> (restype, argtypes). Every symbol declared in include/csd_b200.h must be listed here;
# tests/test_abi.py checks both directions against the header text.
c_float_p = ctypes.POINTER(c_float)
c_int_p = ctypes.POINTER(c_int)

_PROTOTYPES = {
    "csd_last_error": (ctypes.c_char_p, []),
    "csd_abi_version": (c_int, []),
    "csd_device_sm_count": (c_int, [c_int_p]),
    "csd_launch_count": (ctypes.c_longlong, []),
    "csd_upfirdn2d_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64] + [c_int] * 12 + [c_void_p]),
    "csd_upfirdn2d_out_size": (c_int, [c_int] * 12 + [c_int_p, c_int_p]),
    "csd_fused_bias_act_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int64, c_int,
                                       c_int, c_float, c_float, c_void_p]),
    "csd_ve_perturb_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    "csd_sde_perturb_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "csd_inpaint_merge_f32": (c_int, [c_void_p] * 6 + [c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "csd_dsm_loss_f32": (c_int, [c_void_p] * 6 + [c_int, c_int64, c_void_p]),
    "csd_broadcast_table_f32": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "csd_langevin_norms_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "csd_langevin_update_f32": (c_int, [c_void_p] * 6 + [c_int, c_int64, c_float, c_void_p, c_void_p, c_int,
                                        c_void_p]),
    "csd_reverse_diffusion_update_f32": (c_int, [c_void_p] * 5 + [c_int, c_int64, c_void_p, c_void_p, c_int,
                                                 c_void_p, c_int, c_void_p]),
    "csd_euler_maruyama_update_f32": (c_int, [c_void_p] * 5 + [c_int, c_int64, c_void_p, c_void_p, c_float, c_int,
                                              c_void_p, c_int, c_void_p]),
    "csd_step_advance": (c_int, [c_void_p, c_void_p]),
    "csd_nchw_to_nhwc_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                      c_float, c_float, c_void_p]),
    "csd_nhwc_bf16_to_nchw": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p,
                                      c_void_p]),
    "csd_gn_chan_stats_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "csd_gn_finalize_partials_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "csd_gn_coeffs_f32": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                  c_int, c_int, c_float, c_void_p]),
    "csd_gn_coeffs_partials_f32": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                           c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                           c_void_p]),
    "csd_gn_fused_supported": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "csd_gn_fused_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                  c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "csd_gn_apply_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "csd_fir_resample_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                           c_float_p, c_void_p]),
    "csd_fir_norm_resample_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                                c_int, c_int, c_float_p, c_void_p]),
    "csd_fir_norm_resample_nhwc_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                               c_int, c_int, c_float_p, c_void_p]),
    "csd_softmax_rows_f32_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_float, c_void_p]),
    # fp32-activation ("tf32" plan) variants: same argument lists as the *_bf16 entries
    "csd_nchw_to_nhwc_f32": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                     c_float, c_float, c_void_p]),
    "csd_nhwc_f32_to_nchw": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p,
                                     c_void_p]),
    "csd_gn_chan_stats_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "csd_gn_fused_supported_f32": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "csd_gn_fused_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                 c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "csd_gn_apply_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "csd_fir_resample_nhwc_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                          c_float_p, c_void_p]),
    "csd_softmax_rows_f32_f32": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_float, c_void_p]),
    "csd_time_embedding_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    "csd_dense_rows_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "csd_conv_gemm": (c_int, [ctypes.POINTER(ConvGemmDesc), c_void_p]),
    "csd_tap_shift_sum_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                       c_int, c_void_p]),
    "csd_attn_core_supported": (c_int, [c_int, c_int]),
    "csd_attn_core_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                   c_int, c_int, c_int, c_float, c_void_p]),
    "csd_rk_combine_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float_p, c_float, c_void_p, c_void_p]),
    "csd_rk_error_sumsq_f32": (c_int, [c_void_p, c_int64, c_int, c_float_p, c_float, c_void_p, c_void_p, c_float, c_float,
                                       c_void_p, c_void_p]),
    "csd_sumsq_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "csd_fused_adam_ema_f32": (c_int, [c_void_p] * 5 + [c_int64] + [c_float] * 8 + [c_void_p, c_float, c_void_p]),
    "csd_pack_weights": (c_int, [c_void_p, c_int, c_int64, c_void_p]),
    # ---- training backward ----
    "csd_pixmajor_geometry": (c_int, [c_int, c_int, c_int, ctypes.POINTER(PixMajorGeom)]),
    "csd_nhwc_to_pixmajor_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                          ctypes.POINTER(PixMajorGeom), c_int, c_void_p, c_void_p]),
    "csd_wgrad_gemm_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, ctypes.POINTER(PixMajorGeom), c_void_p,
                                    c_void_p]),
    "csd_wgrad_direct_splits": (c_int, [c_int] * 6 + [c_int_p]),
    "csd_wgrad_direct_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_void_p, c_int, c_void_p]),
    "csd_wgrad_reduce_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int64, c_int64, c_int64,
                                     c_int, c_int, c_void_p]),
    "csd_gn_bwd_stats_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                      c_int, c_int, c_int, c_void_p]),
    "csd_gn_bwd_coeffs_f32": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_int, c_float, c_void_p]),
    "csd_gn_bwd_apply_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                      c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "csd_fir_resample_bwd_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float_p, c_int,
                                               c_void_p]),
    "csd_softmax_bwd_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_float,
                                     c_void_p]),
    "csd_transpose_bf16": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_void_p]),
    "csd_axpy_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_int, c_void_p]),
    "csd_dropout_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p, ctypes.c_uint64, c_void_p]),
    "csd_zero_stuff_nhwc_bf16": (c_int, [c_void_p, c_void_p] + [c_int] * 8 + [c_void_p]),
    "csd_nchw_grad_to_nhwc_bf16": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                           c_int, c_int, c_void_p]),
    "csd_bias_temb_grad_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "csd_sgemm_small_f32": (c_int, [c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int, c_void_p, c_int,
                                    c_float, c_void_p, c_int, c_void_p, c_void_p]),
    "csd_silu_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "csd_time_features_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "csd_dsm_loss_bwd_f32": (c_int, [c_void_p] * 7 + [c_int, c_int64, c_void_p]),
}

_lib = None
_lock = threading.Lock()


def lib():
    """Open libcsd_b200.so (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise CsdError(
                    f"{LIB_PATH} is missing: build it with "
                    "`python -m conditional_score_diffusion_b200.build` (needs nvcc). "
                    "There is no CPU or PyTorch fallback for this path."
                )
            handle = ctypes.CDLL(LIB_PATH)
            for name, (restype, argtypes) in _PROTOTYPES.items():
                fn = getattr(handle, name)
                fn.restype = restype
                fn.argtypes = argtypes
            _lib = handle
    return _lib


def check(status):
    if status != 0:
        msg = lib().csd_last_error()
        raise CsdError(f"libcsd_b200 status {status}: {msg.decode() if msg else '?'}")


def exported_names():
    return sorted(_PROTOTYPES)
