"""ctypes loader for libcvae.so.  There is NO fallback: a missing library is a hard error, because
every result this package produces must come from the sm_100a kernels (BASELINE.json north_star)."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libcvae.so")


class CvaeError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make -C critic-vae_b200/csrc` (or __graft_entry__.build()). "
        "There is no CPU or PyTorch fallback for the Critic-VAE hot path.")

lib = ctypes.CDLL(LIB_PATH)

c_int, c_void_p, c_float, c_double, c_i64 = ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_double, ctypes.c_int64


class ConvDesc(ctypes.Structure):
    """cvae_conv_desc (include/cvae.h)."""
    _fields_ = [("batch", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
                ("ksize", ctypes.c_int32), ("src_channels", ctypes.c_int32), ("n_total", ctypes.c_int32),
                ("loader", ctypes.c_int32), ("epilogue", ctypes.c_int32), ("ktab", ctypes.c_int32),
                ("tm", ctypes.c_int32), ("n_block", ctypes.c_int32),
                ("src", c_void_p), ("src2", c_void_p), ("wpack", c_void_p), ("bias", c_void_p),
                ("act", c_void_p), ("out", c_void_p), ("stats", c_void_p),
                ("stack", ctypes.c_int32), ("reserved", ctypes.c_int32), ("workspace", c_void_p), ("workspace_bytes", ctypes.c_int64)]


class WgradDesc(ctypes.Structure):
    """cvae_wgrad_desc (include/cvae.h)."""
    _fields_ = [("kind", ctypes.c_int32), ("batch", ctypes.c_int32), ("height", ctypes.c_int32),
                ("width", ctypes.c_int32), ("cout", ctypes.c_int32), ("cin", ctypes.c_int32),
                ("splits", ctypes.c_int32),
                ("x", c_void_p), ("dy", c_void_p), ("dy2", c_void_p), ("dw", c_void_p),
                ("dbias", c_void_p), ("workspace", c_void_p), ("fold_stream", c_void_p),
                ("workspace_bytes", ctypes.c_int64)]


lib.cvae_last_error.restype = ctypes.c_char_p
lib.cvae_version.restype = c_int
lib.cvae_check_device_fault.argtypes = [c_void_p]
lib.cvae_conv_gemm.argtypes = [ctypes.POINTER(ConvDesc), c_void_p]
lib.cvae_conv_ksteps.argtypes = [c_int, c_int, c_int]
lib.cvae_conv_debug_counters.argtypes = [c_void_p]
lib.cvae_conv_debug_counters.restype = None
lib.cvae_wgrad_debug_counters.argtypes = [c_void_p]
lib.cvae_wgrad_debug_counters.restype = None
lib.cvae_conv_wa_debug_counters.argtypes = [c_void_p]
lib.cvae_conv_wa_debug_counters.restype = None
lib.cvae_conv_wa_tune.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int]
lib.cvae_conv_gemm_workspace_bytes.argtypes = [ctypes.POINTER(ConvDesc)]
lib.cvae_conv_gemm_workspace_bytes.restype = c_i64
lib.cvae_conv_wa_tune.restype = None
lib.cvae_conv_wgrad_workspace_bytes.argtypes = [ctypes.POINTER(WgradDesc)]
lib.cvae_conv_wgrad_workspace_bytes.restype = c_i64
lib.cvae_conv_wgrad.argtypes = [ctypes.POINTER(WgradDesc), c_void_p]


class PackJob(ctypes.Structure):
    """cvae_pack_job (include/cvae.h)."""
    _fields_ = [("kind", ctypes.c_int32), ("n", ctypes.c_int32), ("ksteps", ctypes.c_int32),
                ("k_channels", ctypes.c_int32), ("cout", ctypes.c_int32), ("cin", ctypes.c_int32),
                ("src", c_void_p), ("src2", c_void_p), ("dst", c_void_p)]


P = c_void_p
_SIGS = {
    "cvae_pack_elems": ([ctypes.POINTER(PackJob)], c_i64),
    "cvae_pack_weights": ([ctypes.POINTER(PackJob), c_int, P], c_int),
    "cvae_bn_finalize": ([c_int, c_i64, c_int, P, P, P, P, P, P, P, c_float, c_float, P, P], c_int),
    "cvae_bn_pool_act_fwd": ([c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P], c_int),
    "cvae_bn_fwd": ([c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, c_float, c_float, P, P, P, P, P], c_int),
    "cvae_bn_pool_act_bwd": ([c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P, P], c_int),
    "cvae_fc_fwd": ([c_int, P, P, P, P, P, P], c_int),
    "cvae_fc_bwd": ([c_int, P, P, P, P, P, P, P, P, P], c_int),
    "cvae_decin_fwd": ([c_int, P, P, P, P], c_int),
    "cvae_decin_bwd": ([c_int, P, P, P, P, P, P, P], c_int),
    "cvae_latent_kld_partials": ([c_int], c_int),
    "cvae_latent_fwd": ([c_int, c_int, P, P, P, P, P, P], c_int),
    "cvae_latent_bwd": ([c_int, P, P, P, P, P, c_float, P, P], c_int),
    "cvae_loss_fwd": ([c_int, P, P, P, P, ctypes.POINTER(c_float), c_float, P, P, P, P], c_int),
    "cvae_loss_bwd": ([c_int, P, P, P, ctypes.POINTER(c_float), c_float, P, P, P, P, P, P], c_int),
    "cvae_adam_step": ([c_i64, P, P, P, P, P, c_float, c_float, c_float, c_float, c_float, P], c_int),
    "cvae_adam_update": ([c_i64, P, P, P, P, P, c_float, c_float, c_float, c_float, c_float, c_int, P], c_int),
    "cvae_critic_param_count": ([], c_int),
    "cvae_loss_sums": ([c_int, P, P, P, P, P], c_int),
    "cvae_loss_finalize": ([c_int, P, P, P, ctypes.c_float, P, P, P], c_int),
    "cvae_loss_bwd_sums": ([c_int, P, P, P, P, P, P, P], c_int),
    "cvae_bottleneck_debug": ([P], c_int),
    "cvae_bottleneck_max_clusters": ([c_int, c_i64], c_int),
    "cvae_bottleneck_fwd": ([c_int] + [P] * 11, c_int),
    "cvae_bottleneck_bwd": ([c_int, P, P, P, P, ctypes.c_float, P, P, P, P, P], c_int),
    "cvae_comm_unique_id": ([P], c_int),
    "cvae_comm_init": ([c_int, c_int, P], c_int),
    "cvae_comm_world": ([], c_int),
    "cvae_comm_allreduce_sum": ([P, c_i64, P], c_int),
    "cvae_comm_destroy": ([], c_int),
    "cvae_launch_count": ([], c_i64),
    "cvae_iou_counts": ([c_i64, P, P, P, P], c_int),
    "cvae_frames_u8_to_f32": ([c_int, P, P, P], c_int),
    "cvae_critic_fwd": ([c_int, P, P, P, P], c_int),
    "cvae_diff_grey": ([c_int, P, P, P, P, P], c_int),
    "cvae_mask_iou": ([c_int, P, P, c_double, c_double, c_int, c_int, P, P, P, P, P, P], c_int),
}
for _name, (_args, _res) in _SIGS.items():
    getattr(lib, _name).argtypes = _args
    getattr(lib, _name).restype = _res

EXPORTS = ["cvae_last_error", "cvae_version", "cvae_check_device_fault", "cvae_conv_gemm", "cvae_conv_ksteps",
           "cvae_conv_wgrad_workspace_bytes", "cvae_conv_wgrad", "cvae_conv_debug_counters", "cvae_wgrad_debug_counters",
           "cvae_conv_wa_debug_counters", "cvae_conv_wa_tune", "cvae_conv_gemm_workspace_bytes"] + list(_SIGS)


def check(rc: int) -> None:
    if rc != 0:
        raise CvaeError(f"libcvae error {rc}: {lib.cvae_last_error().decode()}")


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


# enums of include/cvae.h
LOAD_NHWC, LOAD_NCHW3, LOAD_S2D, LOAD_S2D_NCHW3_DTANH = 0, 1, 2, 3
EPI_STATS, EPI_BIAS_RELU, EPI_PHASE_BIAS_RELU, EPI_PHASE_BIAS_TANH, EPI_MASK, EPI_PLAIN = 0, 1, 2, 3, 4, 5
KTAB_GENERIC, KTAB_PAIR8, KTAB_BLOCK64 = 0, 1, 2
PACK_KORDER_BLOCK64, PACK_STACK2, PACK_STACK4, PACK_KORDER_BLOCK32 = 0x100, 0x200, 0x400, 0x800
WGRAD_5X5, WGRAD_PHASE, WGRAD_SHIFT_FRAMES, WGRAD_SHIFT_PHASE12 = 0, 1, 2, 3
PACK_FWD5, PACK_DGRAD5, PACK_PAIR8, PACK_PHASE_FWD, PACK_PHASE_DGRAD, PACK_FC, PACK_DECIN = range(7)
ACT_RELU, ACT_TANH = 0, 1
