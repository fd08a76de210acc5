"""
ctypes binding of libeventflow.so (include/eventflow.h).  There is NO fallback: if the library is missing or a call
fails, an exception is raised -- the product path never computes on the CPU or through stock PyTorch ops.
"""

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libeventflow.so")

EF_LIF, EF_PLIF, EF_ALIF, EF_XLIF = 0, 1, 2, 3
NEURON_CODES = {"lif": EF_LIF, "plif": EF_PLIF, "alif": EF_ALIF, "xlif": EF_XLIF}
SURROGATE_CODES = {"arctanspike": 0, "superspike": 1, "trianglespike": 2, "mgspike": 3}

_f32p = C.c_void_p  # all device pointers travel as void*
_i32 = C.c_int32


class LifConvParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("Cin", _i32), ("C", _i32), ("H", _i32), ("W", _i32),
        ("ksize", _i32), ("stride", _i32), ("neuron", _i32), ("hard_reset", _i32), ("surrogate", _i32),
        ("act_width", C.c_float),
        ("x", _f32p), ("x_cl", _f32p), ("v_in", _f32p), ("z_in", _f32p), ("z_in_cl", _f32p), ("aux_in", _f32p),
        ("w_ff", _f32p), ("w_rec", _f32p), ("leak", _f32p), ("thresh", _f32p), ("leak_aux", _f32p), ("add_pt", _f32p),
        ("t0", _f32p), ("t1", _f32p), ("residual", _f32p), ("w_split", _f32p),
        ("v_out", _f32p), ("z_out", _f32p), ("z_out_cl", _f32p), ("aux_out", _f32p), ("out", _f32p), ("out_cl", _f32p),
    ]  # fmt: skip


class LifConvWindowParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("T", _i32), ("H", _i32), ("W", _i32), ("hard_reset", _i32), ("save_all_v", _i32),
        ("x_cl", _f32p), ("v_in", _f32p), ("z_in_cl", _f32p), ("leak", _f32p), ("thresh", _f32p), ("w_split", _f32p),
        ("v_out", _f32p), ("z_out_cl", _f32p),
    ]  # fmt: skip


class LifConvBwdParams(C.Structure):
    _fields_ = [
        ("f", LifConvParams),
        ("g_out", _f32p), ("g_v_out", _f32p), ("g_z_out", _f32p), ("g_aux_out", _f32p),
        ("scratch_gI", _f32p), ("scratch_gP", _f32p),
        ("g_x", _f32p), ("g_v_in", _f32p), ("g_z_in", _f32p), ("g_aux_in", _f32p),
        ("g_w_ff", _f32p), ("g_w_rec", _f32p), ("g_leak", _f32p), ("g_thresh", _f32p), ("g_leak_aux", _f32p),
        ("g_add_pt", _f32p), ("g_t0", _f32p), ("g_t1", _f32p), ("scratch_gI_up", _f32p), ("scratch_gP_up", _f32p),
        ("reset_grad", _i32), ("neuron_only", _i32),
    ]  # fmt: skip


class Conv32BwdTcParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("H", _i32), ("W", _i32), ("has_rec", _i32),
        ("gI", _f32p), ("x_cl", C.c_void_p), ("z_in_cl", C.c_void_p), ("w_bwd", C.c_void_p), ("gI_hi", C.c_void_p), ("gI_mid", C.c_void_p),
        ("g_x", _f32p), ("g_z_in", _f32p), ("g_z_tmp", _f32p), ("wg_partial", _f32p), ("wg_flags", _i32), ("g_w_ff", _f32p), ("g_w_rec", _f32p),
        ("gP_sum", _f32p), ("x_f32", _f32p),
    ]


class LifBwdTcParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("H", _i32), ("W", _i32), ("has_rec", _i32), ("hard_reset", _i32), ("surrogate", _i32), ("act_width", C.c_float),
        ("x_cl", _f32p), ("z_in_cl", _f32p), ("v_in", _f32p), ("v_out", _f32p), ("g_out", _f32p), ("g_v_out", _f32p), ("g_z_out", _f32p),
        ("leak", _f32p), ("thresh", _f32p), ("w_bwd", _f32p), ("gI_hi", _f32p), ("gI_mid", _f32p),
        ("g_x", _f32p), ("g_v_in", _f32p), ("g_z_in", _f32p), ("g_w_ff", _f32p), ("g_w_rec", _f32p), ("g_leak", _f32p), ("g_thresh", _f32p),
        ("wg_partial", _f32p), ("wg_flags", _i32), ("Cin", _i32), ("x_f32", _f32p), ("gI_f32", _f32p),
    ]  # fmt: skip


EF_WG_ACCUMULATE, EF_WG_FINALIZE = 1, 2
EF_TCG_MAX_SRC = 4


class WSrc(C.Structure):
    _fields_ = [("w", _f32p), ("c_total", _i32), ("ch0", _i32), ("n", _i32), ("split", _i32), ("s2d", _i32)]


class LifConvGParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("H", _i32), ("W", _i32), ("C", _i32), ("n_src", _i32), ("hard_reset", _i32), ("s2d", _i32),
        ("src", _f32p * EF_TCG_MAX_SRC), ("src_c", _i32 * EF_TCG_MAX_SRC),
        ("v_in", _f32p), ("z_in_cl", _f32p), ("residual_cl", _f32p), ("leak", _f32p), ("thresh", _f32p), ("w_image", _f32p),
        ("v_out", _f32p), ("z_out_cl", _f32p), ("out_cl", _f32p),
    ]  # fmt: skip


class LifBwdWindowParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("T", _i32), ("H", _i32), ("W", _i32), ("hard_reset", _i32), ("surrogate", _i32), ("act_width", C.c_float),
        ("x_cl", _f32p), ("z_cl", _f32p), ("z_prev_cl", _f32p), ("v", _f32p), ("v_prev", _f32p), ("g_out", _f32p),
        ("leak", _f32p), ("thresh", _f32p), ("w_bwd", _f32p), ("gI_hi", _f32p), ("gI_mid", _f32p),
        ("g_x", _f32p), ("g_v_prev", _f32p), ("g_w_ff", _f32p), ("g_leak", _f32p), ("g_thresh", _f32p), ("wg_partial", _f32p),
        ("Cin", _i32), ("x_f32", _f32p), ("gI_f32", _f32p),
    ]  # fmt: skip


class PredParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("Cin", _i32), ("Cout", _i32), ("H", _i32), ("W", _i32),
        ("x", _f32p), ("x_cl", _f32p), ("w", _f32p), ("b", _f32p), ("y", _f32p),
        ("g_y", _f32p), ("g_x", _f32p), ("g_w", _f32p), ("g_b", _f32p),
    ]  # fmt: skip


class IweLossParams(C.Structure):
    _fields_ = [
        ("S", _i32), ("B", _i32), ("T", _i32), ("T_maps", _i32), ("H", _i32), ("W", _i32),
        ("n_total", _i32), ("n_per_pass", _i32),
        ("flow_scaling", C.c_float), ("weight", C.c_float),
        ("loss_scaling", _i32), ("smoothing_mask", _i32), ("overwrite_intermediate", _i32),
        ("events", _f32p), ("pol_mask", _f32p), ("flow_maps", _f32p), ("event_mask", _f32p), ("pass_offsets", _f32p),
        ("workspace", _f32p), ("loss", _f32p), ("g_loss", _f32p), ("g_flow_maps", _f32p),
    ]  # fmt: skip


EF_IWE_MAX_PASSES, EF_IWE_MAX_SCALES = 32, 4
EF_HEAD_MAX_CIN = 10


class IweLossPassParams(C.Structure):
    _fields_ = [
        ("S", _i32), ("B", _i32), ("T", _i32), ("T_maps", _i32), ("H", _i32), ("W", _i32),
        ("flow_scaling", C.c_float), ("weight", C.c_float),
        ("loss_scaling", _i32), ("smoothing_mask", _i32), ("overwrite_intermediate", _i32),
        ("n_pass", _i32 * EF_IWE_MAX_PASSES),
        ("events", _f32p * EF_IWE_MAX_PASSES), ("pol_mask", _f32p * EF_IWE_MAX_PASSES),
        ("flow", _f32p * (EF_IWE_MAX_SCALES * EF_IWE_MAX_PASSES)), ("event_mask", _f32p * EF_IWE_MAX_PASSES),
        ("workspace", _f32p), ("loss", _f32p), ("g_loss", _f32p),
        ("g_flow", _f32p * (EF_IWE_MAX_SCALES * EF_IWE_MAX_PASSES)),
    ]  # fmt: skip


class IweImageParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("N", _i32), ("H", _i32), ("W", _i32), ("round_idx", _i32),
        ("tref", C.c_float), ("flow_scaling", C.c_float),
        ("events", _f32p), ("pol_mask", _f32p), ("flow", _f32p), ("event_flow", _f32p), ("iwe", _f32p),
    ]  # fmt: skip


class IweInterpParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("N", _i32), ("H", _i32), ("W", _i32), ("round_idx", _i32),
        ("tref", C.c_float), ("flow_scaling", C.c_float),
        ("events", _f32p), ("flow", _f32p), ("idx", _f32p), ("weights", _f32p),
    ]  # fmt: skip


class ConvAnnParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("C1", _i32), ("C2", _i32), ("Cout", _i32), ("H", _i32), ("W", _i32), ("act", _i32),
        ("x1", _f32p), ("x2", _f32p), ("x2_scale", _f32p),
        ("x1_bstride", C.c_int64), ("x2_bstride", C.c_int64), ("x2_scale_bstride", C.c_int64),
        ("w", _f32p), ("bias", _f32p), ("residual", _f32p), ("blend_h", _f32p), ("blend_u", _f32p),
        ("blend_h_bstride", C.c_int64), ("blend_u_bstride", C.c_int64), ("out", _f32p),
        ("stride", _i32), ("act_out", _f32p), ("inference", _i32),
    ]  # fmt: skip


EF_DP_MAX_RANKS = 8


class DpStepParams(C.Structure):
    _fields_ = [
        ("world", _i32), ("rank", _i32), ("n", _i32), ("step", _i32),
        ("clip", C.c_float), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("bc1", C.c_float), ("bc2_sqrt", C.c_float),
        ("epoch", C.c_uint32), ("epoch_launches", C.c_uint32), ("graceful", _i32), ("grid_expected", _i32),
        ("grads", _f32p * EF_DP_MAX_RANKS), ("signals", _f32p * EF_DP_MAX_RANKS),
        ("param", _f32p), ("m", _f32p), ("v", _f32p), ("sqnorm", _f32p), ("scratch", _f32p), ("status", _f32p),
        ("timeout_ms", _i32),
    ]  # fmt: skip


class AnnGateBwdParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("C", _i32), ("H", _i32), ("W", _i32), ("act", _i32),
        ("g_y", _f32p), ("act_out", _f32p), ("blend_h", _f32p), ("blend_u", _f32p),
        ("blend_h_bstride", C.c_int64), ("blend_u_bstride", C.c_int64),
        ("g_pre", _f32p), ("g_h", _f32p), ("g_u", _f32p), ("g_bias", _f32p),
    ]  # fmt: skip


class IweMetricsParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("T", _i32), ("T_maps", _i32), ("H", _i32), ("W", _i32), ("n_total", _i32), ("n_per_pass", _i32),
        ("flow_scaling", C.c_float),
        ("events", _f32p), ("pol_mask", _f32p), ("flow_maps", _f32p), ("pass_offsets", _f32p), ("workspace", _f32p), ("out", _f32p),
    ]  # fmt: skip


class AeeParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("H", _i32), ("W", _i32), ("flow_scaling", C.c_float),
        ("flow", _f32p), ("gtflow", _f32p), ("event_mask", _f32p), ("dt_ratio", _f32p), ("workspace", _f32p), ("out", _f32p),
    ]  # fmt: skip


class EncodeParams(C.Structure):
    _fields_ = [
        ("B", _i32), ("N", _i32), ("H", _i32), ("W", _i32), ("num_bins", _i32), ("round_ts", _i32),
        ("events", _f32p), ("cnt", _f32p), ("voxel", _f32p), ("mask", _f32p), ("pol_mask", _f32p),
    ]  # fmt: skip


EXPORTS = {
    # name: (restype, argtypes)
    "ef_version": (C.c_int, []),
    "ef_last_error": (C.c_char_p, []),
    "ef_device_ok": (C.c_int, []),
    "ef_launch_count": (C.c_uint64, []),
    "ef_lif_conv_fwd": (C.c_int, [C.POINTER(LifConvParams), C.c_void_p]),
    "ef_lif_conv_fwd_window": (C.c_int, [C.POINTER(LifConvWindowParams), C.c_void_p]),
    "ef_lif_neuron_fwd": (C.c_int, [C.POINTER(LifConvParams), C.c_void_p, C.c_void_p]),
    "ef_lif_conv_bwd": (C.c_int, [C.POINTER(LifConvBwdParams), C.c_void_p]),
    "ef_lif_bwd_tc": (C.c_int, [C.POINTER(LifBwdTcParams), C.c_void_p]),
    "ef_conv32_bwd_tc": (C.c_int, [C.POINTER(Conv32BwdTcParams), C.c_void_p]),
    "ef_split2_pack_cl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, _i32, _i32, C.c_void_p]),
    "ef_wgrad_tcg_partial_elems": (C.c_int64, [_i32, _i32, _i32, _i32, _i32]),
    "ef_wgrad_tcg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, _i32, C.c_void_p, C.c_void_p, _i32, _i32, C.c_void_p]),
    "ef_lif_bwd_window": (C.c_int, [C.POINTER(LifBwdWindowParams), C.c_void_p]),
    "ef_lif_wgrad_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, C.c_void_p, _i32, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "ef_split_weights_g_elems": (C.c_int64, [_i32, _i32, C.POINTER(WSrc)]),
    "ef_split_weights_g": (C.c_int, [C.POINTER(WSrc), _i32, _i32, C.c_void_p, C.c_void_p]),
    "ef_lif_conv_fwd_g": (C.c_int, [C.POINTER(LifConvGParams), C.c_void_p]),
    "ef_split_weights_bwd_elems": (C.c_int64, [_i32]),
    "ef_lif_wgrad_partial_elems": (C.c_int64, [_i32, _i32, _i32, _i32]),
    "ef_split_weights_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ef_split_weights_elems": (C.c_int64, [_i32, _i32, _i32]),
    "ef_split_weights": (C.c_int, [C.c_void_p, C.c_void_p, _i32, _i32, C.c_void_p, C.c_void_p]),
    "ef_debug_tc_trace": (C.c_int, [C.c_void_p]),
    "ef_debug_tc_skip": (C.c_int, [C.c_int]),
    "ef_debug_tc_cpt": (C.c_int, [C.c_int]),
    "ef_debug_pdl": (C.c_int, [C.c_int]),
    "ef_pack_split_cl": (C.c_int, [C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, C.c_void_p]),
    "ef_split_weights_head": (C.c_int, [C.c_void_p, _i32, C.c_void_p, C.c_void_p]),
    "ef_pack_split_s2d_cl": (C.c_int, [C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, C.c_void_p]),
    "ef_space_to_depth_cl": (C.c_int, [C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, C.c_void_p]),
    "ef_pack_cl": (C.c_int, [C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, C.c_void_p]),
    "ef_unpack_cl": (C.c_int, [C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, C.c_void_p]),
    "ef_upsample_bilinear2x": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, _i32, _i32, C.c_void_p]),
    "ef_upsample_bilinear2x_cl": (C.c_int, [C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, C.c_void_p]),
    "ef_upsample_nearest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, _i32, _i32, _i32, _i32, C.c_void_p]),
    "ef_upsample_bilinear2x_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, _i32, _i32, C.c_void_p]),
    "ef_upsample_nearest_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, _i32, _i32, _i32, _i32, C.c_void_p]),
    "ef_pred_fwd": (C.c_int, [C.POINTER(PredParams), C.c_void_p]),
    "ef_pred_bwd": (C.c_int, [C.POINTER(PredParams), C.c_void_p]),
    "ef_iwe_loss_workspace_elems": (C.c_int64, [_i32, _i32, _i32, _i32]),
    "ef_iwe_loss_fwd": (C.c_int, [C.POINTER(IweLossParams), C.c_void_p]),
    "ef_iwe_loss_bwd": (C.c_int, [C.POINTER(IweLossParams), C.c_void_p]),
    "ef_iwe_loss_fwd_passes": (C.c_int, [C.POINTER(IweLossPassParams), C.c_void_p]),
    "ef_iwe_loss_bwd_passes": (C.c_int, [C.POINTER(IweLossPassParams), C.c_void_p]),
    "ef_iwe_image": (C.c_int, [C.POINTER(IweImageParams), C.c_void_p]),
    "ef_iwe_purge_unfeasible": (C.c_int, [C.c_void_p, C.c_int64, _i32, _i32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ef_iwe_get_interpolation": (C.c_int, [C.POINTER(IweInterpParams), C.c_void_p]),
    "ef_iwe_interpolate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, C.c_void_p, C.c_void_p]),
    "ef_spike_fwd": (C.c_int, [C.c_void_p, C.c_void_p, _i32, _i32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "ef_spike_bwd": (C.c_int, [C.c_void_p, C.c_void_p, _i32, _i32, C.c_int64, C.c_int64, C.c_void_p, _i32, C.c_float, C.c_void_p, C.c_void_p]),
    "ef_conv_ann_fwd": (C.c_int, [C.POINTER(ConvAnnParams), C.c_void_p]),
    "ef_conv3x3_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, _i32, C.c_void_p]),
    "ef_conv3x3_bwd_s": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, _i32, _i32, C.c_void_p]),
    "ef_ann_gate_bwd": (C.c_int, [C.POINTER(AnnGateBwdParams), C.c_void_p]),
    "ef_dp_step": (C.c_int, [C.POINTER(DpStepParams), C.c_void_p]),
    "ef_dp_step_grid": (_i32, [_i32]),
    "ef_ipc_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]),
    "ef_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "ef_ipc_close": (C.c_int, [C.c_void_p]),
    "ef_ipc_free": (C.c_int, [C.c_void_p]),
    "ef_ann_cat_scale": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, _i32, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "ef_ann_scale_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _i32, _i32, _i32, _i32, _i32, C.c_int64, C.c_int64, C.c_void_p]),
    "ef_iwe_metrics_workspace_elems": (C.c_int64, [_i32, _i32, _i32]),
    "ef_iwe_metrics": (C.c_int, [C.POINTER(IweMetricsParams), C.c_void_p]),
    "ef_aee": (C.c_int, [C.POINTER(AeeParams), C.c_void_p]),
    "ef_encode_events": (C.c_int, [C.POINTER(EncodeParams), C.c_void_p]),
    "ef_grad_sqnorm": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ef_clip_adam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_float,
                               C.c_float, C.c_float, C.c_float, _i32, C.c_void_p]),
}  # fmt: skip

_lib = None
LAUNCHES = 0  # number of library compute calls issued by this process
GRAPH_KERNELS = 0  # kernels executed through CUDA-graph replays (not seen by ef_launch_count)


class EventFlowError(RuntimeError):
    pass


def lib():
    """Load libeventflow.so once.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EventFlowError(
                f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). There is no CPU / PyTorch fallback."
            )
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().ef_last_error().decode(errors="replace")
        raise EventFlowError(f"{what} failed (rc={rc}): {msg}")


def planes(t):
    """Device pointers of the slices t[0], t[1], ... of a contiguous CUDA tensor (no view tensors are created)."""
    base, step = ptr(t), t.stride(0) * t.element_size()
    return [base + i * step for i in range(t.shape[0])]


def ptr(t):
    """Device pointer of a tensor (or None).  Tensors must be contiguous CUDA tensors."""
    if t is None:
        return None
    if not t.is_cuda:
        raise EventFlowError("event_flow_b200 kernels need CUDA tensors; got a %s tensor (no CPU fallback)" % t.device)
    if not t.is_contiguous():
        raise EventFlowError("event_flow_b200 kernels need contiguous tensors")
    return t.data_ptr()


def stream():
    """Raw handle of torch's current stream on the current device (the C-level getter: the Stream object costs several microseconds)."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


PROFILE = None  # set to a list to record (name, tag, start_event, end_event) around every struct-taking call (bench.py roofline)


def call(name, params, tag=None):
    """Invoke a struct-taking entry point on the current torch stream."""
    global LAUNCHES
    LAUNCHES += 1
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(getattr(lib(), name)(C.byref(params), stream()), name)
        e1.record()
        PROFILE.append((name, tag, e0, e1))
        return
    check(getattr(lib(), name)(C.byref(params), stream()), name)
