"""ctypes binding of the deepsee_b200 C ABI (include/deepsee_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``build.sh`` into
``deepsee_b200/lib/libdeepsee_b200.so``.  There is no fallback: if the library is missing,
importing a compute op raises, and every entry point fails on a non-sm_100 device.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DSEE_LIB_PATH") or os.path.join(_HERE, "lib", "libdeepsee_b200.so")

# every symbol include/deepsee_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "dsee_version", "dsee_last_error", "dsee_launch_count",
    "dsee_peer_exchange_bytes", "dsee_peer_allreduce_small", "dsee_labels_u8", "dsee_bicubic_clamp", "dsee_maxpool2_fwd", "dsee_maxpool2_bwd", "dsee_noise_fill", "dsee_noise_epoch_advance", "dsee_noise_epoch_set", "dsee_onehot_from_labels", "dsee_labels_from_onehot", "dsee_resize_labels",
    "dsee_shared_mlp_fwd", "dsee_style_gather_fwd",
    "dsee_prep_conv_weight", "dsee_prep_conv_weight_f8", "dsee_prep_mod_weight_batched", "dsee_split_f16",
    "dsee_conv3x3_wgrad_per_image_workspace_floats", "dsee_conv3x3_wgrad2_per_image",
    "dsee_subpixel_wgrad_workspace_floats", "dsee_subpixel_wgrad", "dsee_subpixel_dgrad", "dsee_prep_conv_weight_ex", "dsee_split_f16_ups2",
    "dsee_fold2x2", "dsee_conv2d_tc", "dsee_conv2d_tc_wgrad_workspace_floats", "dsee_conv2d_tc_wgrad",
    "dsee_head_gather_fwd", "dsee_head_scatter_bwd", "dsee_conv3x3_fwd", "dsee_conv_pair_mode", "dsee_conv3x3_stats_tiles", "dsee_spade_modulate_fwd",
    "dsee_spade_modulate_bwd", "dsee_spade_modulate_bwd_saved", "dsee_dgrad_modulate_bwd", "dsee_grad_prep_blocks", "dsee_grad_prep", "dsee_reduce_partials",
    "dsee_conv3x3_wgrad_workspace_floats", "dsee_conv3x3_wgrad", "dsee_conv3x3_wgrad2", "dsee_bn_bwd_blocks", "dsee_bn_bwd",
    "dsee_actv_grad_prep", "dsee_onehot_planes", "dsee_shared_mlp_bwd_blocks", "dsee_shared_mlp_bwd", "dsee_style_gather_bwd",
    "dsee_stem_bwd_blocks", "dsee_stem_bwd", "dsee_head_bwd_blocks", "dsee_head_bwd",
    "dsee_bn_stats", "dsee_bn_finalize", "dsee_bn_eval_affine",
    "dsee_stem_fwd", "dsee_head_fwd",
    "dsee_conv2d_direct_fwd", "dsee_instance_norm_fwd", "dsee_instance_norm_workspace_bytes", "dsee_region_pool_chunks",
    "dsee_region_pool_fwd", "dsee_nchw_to_nhwc", "dsee_disc_input", "dsee_avgpool3s2_fwd",
    "dsee_spectral_workspace_floats", "dsee_spectral_weight_fwd", "dsee_spectral_weight_bwd",
    "dsee_spectral_weight_fwd_batched",
    "dsee_modweight_fwd", "dsee_modweight_bwd_workspace_bytes", "dsee_modweight_bwd",
    "dsee_act_bwd", "dsee_conv2d_direct_dgrad", "dsee_conv2d_direct_wgrad_workspace_floats",
    "dsee_conv2d_direct_wgrad", "dsee_channel_sum_chunks", "dsee_channel_sum",
    "dsee_instance_norm_bwd", "dsee_region_pool_bwd", "dsee_avgpool3s2_bwd", "dsee_disc_input_bwd",
]


class ConvOperands(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("a_hi", C.c_void_p * 2), ("a_lo", C.c_void_p * 2), ("a_channels", C.c_int * 2),
        ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("w_inv_scale", C.c_void_p),
        ("n_total", C.c_int), ("passes", C.c_int), ("a_dtype", C.c_int), ("w_dtype", C.c_int),
        ("a_inv_scale", C.c_void_p),
        ("a8_lo", C.c_void_p), ("a8_hi", C.c_void_p), ("w8", C.c_void_p),
        ("w_batch_rows", C.c_int), ("a_sub", C.c_int), ("sub_py", C.c_int), ("sub_px", C.c_int),
    ]


class ConvEpilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("res_ups", C.c_int),
        ("noise", C.c_void_p * 2), ("noise_w", C.c_void_p * 2),
        ("out", C.c_void_p), ("stats_partial", C.c_void_p), ("act_mask", C.c_void_p),
        ("amax_out", C.c_void_p), ("lrelu", C.c_int), ("noise_seed", C.c_uint64 * 2),
        ("act16_hi", C.c_void_p), ("act16_lo", C.c_void_p),
    ]


class ModulateArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_ups", C.c_int),
        ("noise", C.c_void_p), ("noise_w", C.c_void_p),
        ("bn_scale", C.c_void_p), ("bn_shift", C.c_void_p),
        ("gamma_bias", C.c_void_p), ("beta_bias", C.c_void_p),
        ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
        ("C", C.c_int), ("g_hi", C.c_void_p), ("g_lo", C.c_void_p), ("noise_seed", C.c_uint64),
        ("out8_lo", C.c_void_p), ("out8_hi", C.c_void_p),
    ]


class Conv2dTCArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("Hi", C.c_int), ("Wi", C.c_int),
        ("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("Ci", C.c_int), ("a_inv_scale", C.c_void_p),
        ("KH", C.c_int), ("KW", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("w_inv_scale", C.c_void_p),
        ("n_total", C.c_int), ("passes", C.c_int), ("transposed", C.c_int),
        ("Ho", C.c_int), ("Wo", C.c_int),
    ]


class ModulateBwdArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_ups", C.c_int),
        ("noise", C.c_void_p), ("noise_w", C.c_void_p),
        ("bn_scale", C.c_void_p), ("bn_shift", C.c_void_p),
        ("gamma_bias", C.c_void_p), ("dt", C.c_void_p),
        ("dxhat", C.c_void_p), ("dgb_hi", C.c_void_p), ("dgb_lo", C.c_void_p),
        ("partial", C.c_void_p), ("C", C.c_int),
        ("dt_amax", C.c_void_p), ("dgb_inv_scale", C.c_void_p),
    ]


class DgradModBwdArgs(C.Structure):
    _fields_ = [
        ("act_mask", C.c_void_p), ("g_hi", C.c_void_p), ("g_lo", C.c_void_p),
        ("x", C.c_void_p), ("x_ups", C.c_int),
        ("noise", C.c_void_p), ("noise_seed", C.c_uint64), ("noise_w", C.c_void_p),
        ("bn_scale", C.c_void_p), ("bn_shift", C.c_void_p),
        ("dy_amax", C.c_void_p), ("w_l1", C.c_void_p),
        ("dxhat", C.c_void_p), ("dgb_hi", C.c_void_p), ("dgb_lo", C.c_void_p),
        ("dgb_inv_scale", C.c_void_p), ("partial", C.c_void_p), ("C", C.c_int),
    ]


class SnItem(C.Structure):
    _fields_ = [
        ("w_orig", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("w_eff", C.c_void_p),
        ("sigma2", C.c_void_p), ("workspace", C.c_void_p), ("u_saved", C.c_void_p), ("v_saved", C.c_void_p),
        ("N", C.c_int), ("K", C.c_int),
    ]


class ModWeightArgs(C.Structure):
    _fields_ = [
        ("w_seg", C.c_void_p * 2), ("w_sty", C.c_void_p * 2), ("b_seg", C.c_void_p * 2),
        ("b_sty", C.c_void_p * 2), ("alpha", C.c_void_p * 2),
        ("C", C.c_int), ("c1", C.c_int), ("c2", C.c_int), ("plus_one", C.c_int),
    ]


class ModWeightGrads(C.Structure):
    _fields_ = [
        ("dw_seg", C.c_void_p * 2), ("dw_sty", C.c_void_p * 2), ("db_seg", C.c_void_p * 2),
        ("db_sty", C.c_void_p * 2), ("dalpha", C.c_void_p * 2),
    ]


ABI_VERSION = 4
_lib = None


def load():
    """Loads the shared library (once). Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "deepsee_b200: %s not found - run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or ./build.sh). There is no CPU / PyTorch fallback for the hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.dsee_version.restype = C.c_int
    lib.dsee_last_error.restype = C.c_char_p
    lib.dsee_launch_count.restype = C.c_int64
    vp, i, f, d, i64, u64 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_int64, C.c_uint64
    sig = {
        "dsee_onehot_from_labels": [vp, vp, i, i, i, i, vp, vp],
        "dsee_labels_from_onehot": [vp, vp, i, i, i, i, vp, vp],
        "dsee_resize_labels": [vp, vp, i, i, i, i, i, vp],
        "dsee_shared_mlp_fwd": [vp, vp, vp, vp, vp, i, i, i, i, i, i, vp, vp],
        "dsee_style_gather_fwd": [vp, vp, vp, vp, i, i, i, i, i, vp],
        "dsee_prep_conv_weight": [vp, vp, vp, vp, i, i, i, vp],
        "dsee_prep_conv_weight_f8": [vp, vp, vp, i, i, vp],
        "dsee_prep_mod_weight_batched": [vp, vp, vp, vp, vp, i, i, i, i, i, vp],
        "dsee_conv3x3_wgrad2_per_image": [vp, vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i), i, i, i, i,
                                          i, i, vp, vp, vp],
        "dsee_subpixel_wgrad": [vp, vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i), i, i, i, i, i, i, i,
                                vp, vp, vp],
        "dsee_subpixel_dgrad": [vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, vp, vp, vp],
        "dsee_split_f16": [vp, vp, vp, i64, vp],
        "dsee_prep_conv_weight_ex": [vp, vp, vp, vp, i, i, i, i, i, vp],
        "dsee_split_f16_ups2": [vp, vp, vp, i, i, i, i, vp],
        "dsee_fold2x2": [vp, vp, i, i, i, i, vp],
        "dsee_conv2d_tc": [C.POINTER(Conv2dTCArgs), C.POINTER(ConvEpilogue), vp],
        "dsee_conv2d_tc_wgrad": [vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, i, i, i, vp, vp, vp],
        "dsee_conv3x3_fwd": [C.POINTER(ConvOperands), C.POINTER(ConvEpilogue), vp],
        "dsee_conv_pair_mode": [i],
        "dsee_head_gather_fwd": [vp, vp, vp, i, i, i, vp],
        "dsee_head_scatter_bwd": [vp, vp, vp, i, i, i, vp],
        "dsee_conv3x3_stats_tiles": [i, i, i],
        "dsee_spade_modulate_fwd": [C.POINTER(ConvOperands), C.POINTER(ModulateArgs), vp],
        "dsee_spade_modulate_bwd": [C.POINTER(ConvOperands), C.POINTER(ModulateBwdArgs), vp],
        "dsee_dgrad_modulate_bwd": [C.POINTER(ConvOperands), C.POINTER(DgradModBwdArgs), vp],
        "dsee_spade_modulate_bwd_saved": [vp, i, vp, u64, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, vp, vp,
                                          vp, vp, vp, vp],
        "dsee_noise_fill": [u64, vp, i64, vp],
        "dsee_noise_epoch_advance": [vp],
        "dsee_labels_u8": [vp, vp, i64, i, vp, vp],
        "dsee_peer_allreduce_small": [C.POINTER(vp), i, i, i, i, vp, vp, i, vp],
        "dsee_maxpool2_fwd": [vp, vp, i, i, i, i, vp],
        "dsee_maxpool2_bwd": [vp, vp, vp, i, i, i, i, vp],
        "dsee_bicubic_clamp": [vp, vp, i, i, i, i, i, i, vp],
        "dsee_noise_epoch_set": [u64, vp],
        "dsee_grad_prep_blocks": [i64],
        "dsee_grad_prep": [vp, vp, vp, vp, vp, vp, u64, u64, i64, i, vp, vp, vp],
        "dsee_reduce_partials": [vp, i, i, i, f, vp, vp],
        "dsee_conv3x3_wgrad": [vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, vp, vp, i, vp],
        "dsee_conv3x3_wgrad2": [vp, vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i), i, i, i, i, i, i,
                                vp, vp, i, vp],
        "dsee_bn_bwd_blocks": [i, i, i],
        "dsee_bn_bwd": [vp, vp, i, vp, u64, vp, vp, vp, vp, f, vp, i, i, i, i, vp, vp, i, vp, vp],
        "dsee_actv_grad_prep": [vp, i, i, vp, vp, i, i, i, i, i, vp, vp, vp, vp, vp],
        "dsee_onehot_planes": [vp, vp, i64, i, vp],
        "dsee_shared_mlp_bwd_blocks": [i, i, i],
        "dsee_shared_mlp_bwd": [vp, i, i, vp, vp, i, i, i, i, i, i, vp, vp, vp],
        "dsee_style_gather_bwd": [vp, i, i, vp, vp, vp, i, i, i, i, vp],
        "dsee_stem_bwd_blocks": [i, i, i],
        "dsee_stem_bwd": [vp, vp, i, i, i, i, vp, vp, vp],
        "dsee_head_bwd_blocks": [i, i, i],
        "dsee_head_bwd": [vp, vp, vp, vp, i, i, i, i, vp, vp, vp, vp],
        "dsee_bn_stats": [vp, i, vp, u64, vp, i, i, i, i, vp, C.POINTER(C.c_int), vp],
        "dsee_bn_finalize": [vp, i, i, d, d, f, f, vp, vp, vp, vp, vp, vp, vp],
        "dsee_bn_eval_affine": [vp, vp, f, i, vp, vp, vp],
        "dsee_stem_fwd": [vp, vp, vp, vp, i, i, i, i, vp],
        "dsee_head_fwd": [vp, vp, vp, vp, i, i, i, i, vp],
        "dsee_conv2d_direct_fwd": [vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, i, i, vp],
        "dsee_instance_norm_fwd": [vp, vp, vp, vp, vp, i, i, i, f, i, vp],
        "dsee_region_pool_chunks": [i],
        "dsee_region_pool_fwd": [vp, vp, vp, vp, i, i, i, i, vp],
        "dsee_nchw_to_nhwc": [vp, vp, i, i, i, i, i, vp],
        "dsee_disc_input": [vp, vp, vp, vp, i, i, i, i, i, vp],
        "dsee_avgpool3s2_fwd": [vp, vp, i, i, i, i, vp],
        "dsee_spectral_weight_fwd": [vp, vp, vp, i, i, i, f, vp, vp, vp, vp],
        "dsee_spectral_weight_fwd_batched": [C.POINTER(SnItem), i, i, f, vp],
        "dsee_spectral_weight_bwd": [vp, vp, vp, vp, vp, i, i, vp, vp, vp],
        "dsee_modweight_fwd": [C.POINTER(ModWeightArgs), vp, vp, vp, vp],
        "dsee_modweight_bwd": [C.POINTER(ModWeightArgs), vp, vp, vp, C.POINTER(ModWeightGrads), vp, vp],
        "dsee_act_bwd": [vp, vp, vp, i64, i, vp],
        "dsee_conv2d_direct_dgrad": [vp, vp, vp, i, i, i, i, i, i, i, i, i, i, vp],
        "dsee_conv2d_direct_wgrad": [vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, i, vp],
        "dsee_channel_sum_chunks": [i64],
        "dsee_channel_sum": [vp, i64, i, vp, vp, vp],
        "dsee_instance_norm_bwd": [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, vp],
        "dsee_region_pool_bwd": [vp, vp, vp, i, i, i, i, vp],
        "dsee_avgpool3s2_bwd": [vp, vp, i, i, i, i, vp],
        "dsee_disc_input_bwd": [vp, vp, i, i, i, i, i, vp],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.dsee_conv3x3_wgrad_workspace_floats.argtypes = [i, i, i, i, i]
    lib.dsee_conv3x3_wgrad_workspace_floats.restype = C.c_int64
    lib.dsee_conv3x3_wgrad_per_image_workspace_floats.argtypes = [i, i, i, i, i]
    lib.dsee_conv3x3_wgrad_per_image_workspace_floats.restype = C.c_int64
    lib.dsee_subpixel_wgrad_workspace_floats.argtypes = [i, i, i, i, i]
    lib.dsee_subpixel_wgrad_workspace_floats.restype = C.c_int64
    lib.dsee_peer_exchange_bytes.argtypes = [i, i, i]
    lib.dsee_peer_exchange_bytes.restype = C.c_int64
    lib.dsee_spectral_workspace_floats.argtypes = [i, i]
    lib.dsee_spectral_workspace_floats.restype = C.c_int64
    lib.dsee_modweight_bwd_workspace_bytes.argtypes = [i, i, i]
    lib.dsee_modweight_bwd_workspace_bytes.restype = C.c_int64
    lib.dsee_instance_norm_workspace_bytes.argtypes = [i, i, i]
    lib.dsee_instance_norm_workspace_bytes.restype = C.c_int64
    lib.dsee_conv2d_tc_wgrad_workspace_floats.argtypes = [i, i, i, i, i, i, i]
    lib.dsee_conv2d_tc_wgrad_workspace_floats.restype = C.c_int64
    lib.dsee_conv2d_direct_wgrad_workspace_floats.argtypes = [i, i, i, i, i, i, i]
    lib.dsee_conv2d_direct_wgrad_workspace_floats.restype = C.c_int64
    if lib.dsee_version() != ABI_VERSION:
        raise RuntimeError("deepsee_b200: ABI version mismatch")
    if os.environ.get("DSEE_CTA_PAIR", "0") == "1":   # opt-in cta_group::2 main convs
        lib.dsee_conv_pair_mode(1)
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().dsee_last_error()
        raise RuntimeError("deepsee_b200 C-ABI call failed (rc=%d): %s" %
                           (rc, msg.decode() if msg else "?"))


def launch_count():
    return int(load().dsee_launch_count())
