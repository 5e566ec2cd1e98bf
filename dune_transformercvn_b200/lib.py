"""ctypes binding of libtcvn.so (the C ABI in include/tcvn.h).

There is no fallback: if the shared library is missing or a call fails, this raises.  The
library is built in-tree by ``__graft_entry__.build()`` / ``make -C dune_transformercvn_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtcvn.so")

TCVN_FP32, TCVN_BF16 = 0, 1
TCVN_VAL_F32, TCVN_VAL_U8 = 0, 1
TCVN_NCHW_F32 = 0
SEQ_TOKENS, SEQ_ENCODER, SEQ_HEADS = 1, 2, 4
MAX_BLOCKS, MAX_DECODER_LAYERS = 8, 8

# every symbol include/tcvn.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = (
    "tcvn_abi_version", "tcvn_last_error", "tcvn_launch_count", "tcvn_densify", "tcvn_densify_noise",
    "tcvn_collate_workspace_bytes", "tcvn_collate_coords",
    "tcvn_cnn_arena_floats", "tcvn_cnn_packed_bytes", "tcvn_cnn_pack", "tcvn_cnn_workspace_bytes",
    "tcvn_cnn_forward", "tcvn_cnn_forward_sparse", "tcvn_cnn_workspace_bytes_sparse", "tcvn_cnn_run_layer", "tcvn_cnn_read_stage",
    "tcvn_seq_packed_bytes", "tcvn_seq_pack", "tcvn_seq_workspace_bytes", "tcvn_seq_forward",
    "tcvn_t_gemm", "tcvn_t_wgrad", "tcvn_t_colsums", "tcvn_t_bn_finalize", "tcvn_t_bnact_bwd_apply", "tcvn_t_add_colsums",
    "tcvn_t_bnact_fwd", "tcvn_t_pool", "tcvn_t_dropout", "tcvn_t_stem_conv", "tcvn_t_layernorm", "tcvn_t_attention",
    "tcvn_t_eltwise", "tcvn_t_tokens", "tcvn_t_act_pool2", "tcvn_t_act_gap", "tcvn_adamw_workspace_bytes", "tcvn_adamw_fused", "tcvn_set_seed_offset",
    "tcvn_cnn_train_workspace_bytes", "tcvn_cnn_train_forward", "tcvn_cnn_train_backward",
    "tcvn_seq_train_workspace_bytes", "tcvn_seq_train_forward", "tcvn_seq_train_backward",
    "tcvn_t_umma_wgrad", "tcvn_t_umma_wgrad_workspace_bytes", "tcvn_t_umma_conv2_dgrad",
    "tcvn_loss_forward", "tcvn_loss_backward", "tcvn_metrics_update",
    "tcvn_sdxl_pixels_to_ring", "tcvn_sdxl_groupnorm", "tcvn_sdxl_patch_s2", "tcvn_set_sm_limit",
    "tcvn_sdxl16_patch27", "tcvn_sdxl16_groupnorm_workspace_bytes", "tcvn_sdxl16_groupnorm", "tcvn_sdxl16_patch_s2", "tcvn_sdxl16_s2d", "tcvn_sdxl16_conv_s2",
    "tcvn_sdxl16_to_f32", "tcvn_sdxl16_conv", "tcvn_sdxl16_conv2d_stat_bytes", "tcvn_sdxl16_conv2d_c64", "tcvn_sdxl16_gn_stats",
)


class CnnDesc(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("init_features", C.c_int32), ("growth", C.c_int32),
                ("bn_size", C.c_int32), ("num_blocks", C.c_int32), ("block_layers", C.c_int32 * MAX_BLOCKS),
                ("out_features", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("bn_eps", C.c_float)]


class SeqDesc(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("heads", C.c_int32), ("layers", C.c_int32), ("ffn", C.c_int32),
                ("pixel_dim", C.c_int32), ("feature_dim", C.c_int32), ("position_dim", C.c_int32),
                ("num_event_classes", C.c_int32), ("num_prong_classes", C.c_int32),
                ("num_decoder_layers", C.c_int32), ("decoder_widths", C.c_int32 * MAX_DECODER_LAYERS),
                ("bn_eps", C.c_float), ("ln_eps", C.c_float)]


class TcvnError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TcvnError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float
    lib.tcvn_abi_version.restype = C.c_int
    lib.tcvn_last_error.restype = C.c_char_p
    lib.tcvn_launch_count.restype = C.c_longlong
    lib.tcvn_densify.argtypes = [vp, vp, i32, i64, i32, i32, i32, i32, f32, vp, i32, vp]
    lib.tcvn_densify_noise.argtypes = [vp, vp, i32, i64, i32, i32, i32, i32, f32, f32, C.c_uint64, vp, i32, vp]
    lib.tcvn_collate_workspace_bytes.argtypes = [i32]
    lib.tcvn_collate_workspace_bytes.restype = sz
    lib.tcvn_collate_coords.argtypes = [vp, i64, vp, vp, vp, i32, i32, vp, vp, sz, vp]
    lib.tcvn_cnn_arena_floats.argtypes = [C.POINTER(CnnDesc)]
    lib.tcvn_cnn_arena_floats.restype = i64
    lib.tcvn_cnn_packed_bytes.argtypes = [C.POINTER(CnnDesc), i32]
    lib.tcvn_cnn_packed_bytes.restype = sz
    lib.tcvn_cnn_pack.argtypes = [C.POINTER(CnnDesc), i32, vp, vp, sz, vp]
    lib.tcvn_cnn_workspace_bytes.argtypes = [C.POINTER(CnnDesc), i32, i32]
    lib.tcvn_cnn_workspace_bytes.restype = sz
    lib.tcvn_cnn_workspace_bytes_sparse.argtypes = [C.POINTER(CnnDesc), i32, i32, i64]
    lib.tcvn_cnn_workspace_bytes_sparse.restype = sz
    lib.tcvn_cnn_forward.argtypes = [C.POINTER(CnnDesc), i32, vp, vp, i32, vp, vp, sz, vp]
    lib.tcvn_cnn_forward_sparse.argtypes = [C.POINTER(CnnDesc), i32, vp, vp, vp, i32, i64, f32, i32, vp, vp, sz, vp]
    lib.tcvn_cnn_run_layer.argtypes = [C.POINTER(CnnDesc), i32, vp, vp, sz, i32, i32, i32, i32, vp]
    lib.tcvn_cnn_read_stage.argtypes = [C.POINTER(CnnDesc), i32, vp, i32, i32, vp, C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int32), vp]
    lib.tcvn_seq_packed_bytes.argtypes = [C.POINTER(SeqDesc)]
    lib.tcvn_seq_packed_bytes.restype = sz
    lib.tcvn_seq_pack.argtypes = [C.POINTER(SeqDesc), vp, vp, vp, vp, vp, vp, sz, vp]
    lib.tcvn_seq_workspace_bytes.argtypes = [C.POINTER(SeqDesc), i32, i32]
    lib.tcvn_seq_workspace_bytes.restype = sz
    lib.tcvn_seq_forward.argtypes = [C.POINTER(SeqDesc), vp, i32, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, sz, vp]
    u64, f64 = C.c_uint64, C.c_double
    lib.tcvn_t_gemm.argtypes = [vp, i32, i64, i32, i32, vp, vp, i32, vp, i32, i32, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.tcvn_t_wgrad.argtypes = [vp, i32, i64, i32, i32, vp, vp, i32, i32, vp, i32, i32, i32, i32, i32, vp, vp]
    lib.tcvn_t_colsums.argtypes = [i32, vp, i32, i32, vp, i32, i32, vp, i32, i64, i32, i32, vp, vp]
    lib.tcvn_t_bn_finalize.argtypes = [vp, i32, f64, vp, vp, vp, f32, f32, vp, vp, vp, vp]
    lib.tcvn_t_bnact_bwd_apply.argtypes = [vp, i32, i32, vp, i32, i32, vp, vp, i32, f64, vp, i32, i32, i32, i64, i32, i32,
                                           vp, vp, vp, vp]
    lib.tcvn_t_add_colsums.argtypes = [vp, i32, vp, vp]
    lib.tcvn_t_bnact_fwd.argtypes = [vp, i32, i32, vp, i32, i64, i32, i32, vp, i32, i32, vp]
    lib.tcvn_t_pool.argtypes = [i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.tcvn_t_dropout.argtypes = [vp, i32, i32, i32, i64, u64, u64, f32, vp]
    lib.tcvn_t_stem_conv.argtypes = [vp, i32, i32, i32, i32, vp, vp, i32, vp, vp, vp, vp]
    lib.tcvn_t_layernorm.argtypes = [i32, vp, vp, i32, i32, vp, vp, f32, vp, vp, vp, vp, vp, vp]
    lib.tcvn_t_attention.argtypes = [i32, vp, vp, i32, i32, i32, i32, vp, vp, vp, f32, u64, u64, vp]
    lib.tcvn_t_eltwise.argtypes = [i32, vp, vp, vp, i32, i64, vp, vp]
    lib.tcvn_t_tokens.argtypes = [i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.tcvn_t_act_pool2.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, i32, i32, vp]
    lib.tcvn_t_act_gap.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.tcvn_adamw_workspace_bytes.argtypes = []
    lib.tcvn_adamw_workspace_bytes.restype = sz
    lib.tcvn_adamw_fused.argtypes = [vp, vp, vp, vp, i64, vp, i32, vp, vp, vp, vp, vp, vp, f32, f32, vp, sz, vp, vp, vp]
    lib.tcvn_set_seed_offset.argtypes = [vp]
    lib.tcvn_t_umma_wgrad.argtypes = [vp, i64, i32, i32, i32, vp, vp, vp, vp, i32, vp, i32, i32, i32, vp, vp, sz, vp]
    lib.tcvn_t_umma_wgrad_workspace_bytes.argtypes = [i32]
    lib.tcvn_t_umma_wgrad_workspace_bytes.restype = sz
    lib.tcvn_t_umma_conv2_dgrad.argtypes = [vp, vp, i64, i32, i32, vp, vp]
    lib.tcvn_cnn_train_workspace_bytes.argtypes = [C.POINTER(CnnDesc), i32, i32]
    lib.tcvn_cnn_train_workspace_bytes.restype = sz
    lib.tcvn_cnn_train_forward.argtypes = [C.POINTER(CnnDesc), i32, vp, vp, i32, f32, f32, u64, u64, vp, vp, sz, vp]
    lib.tcvn_cnn_train_backward.argtypes = [C.POINTER(CnnDesc), i32, vp, vp, vp, i32, f32, u64, u64, vp, vp, sz, vp]
    lib.tcvn_seq_train_workspace_bytes.argtypes = [C.POINTER(SeqDesc), i32, i32, i32]
    lib.tcvn_seq_train_workspace_bytes.restype = sz
    lib.tcvn_seq_train_forward.argtypes = [C.POINTER(SeqDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, f32, u64,
                                           vp, vp, vp, sz, vp]
    lib.tcvn_seq_train_backward.argtypes = [C.POINTER(SeqDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32,
                                            u64, vp, vp, vp, vp, vp, sz, vp]
    lib.tcvn_loss_forward.argtypes = [vp, vp, i32, i32, vp, vp, i32, i32, i64, i64, f32, f32, f32, vp, vp, vp, vp]
    lib.tcvn_loss_backward.argtypes = [vp, vp, i64, vp, i64, vp, vp, vp]
    lib.tcvn_metrics_update.argtypes = [vp, vp, i32, i32, vp, vp, i32, i32, i64, i64, vp, vp, vp, vp]
    lib.tcvn_sdxl_pixels_to_ring.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    lib.tcvn_sdxl_groupnorm.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, f32, i32, vp, vp, vp]
    lib.tcvn_sdxl_patch_s2.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    lib.tcvn_set_sm_limit.argtypes = [i32]
    lib.tcvn_sdxl16_patch27.argtypes = [vp, i32, i32, i32, i32, f32, vp, vp]
    lib.tcvn_sdxl16_groupnorm_workspace_bytes.argtypes = [i32]
    lib.tcvn_sdxl16_groupnorm_workspace_bytes.restype = sz
    lib.tcvn_sdxl16_groupnorm.argtypes = [vp, i32, i32, i32, i32, vp, vp, f32, i32, vp, vp, sz, vp]
    lib.tcvn_sdxl16_patch_s2.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    lib.tcvn_sdxl16_s2d.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    lib.tcvn_sdxl16_conv_s2.argtypes = [vp, i64, i32, vp, i32, vp, vp, vp, i32, i32, i32, vp]
    lib.tcvn_sdxl16_to_f32.argtypes = [vp, i64, vp, vp]
    lib.tcvn_sdxl16_conv2d_stat_bytes.argtypes = [i32, i32, i32]
    lib.tcvn_sdxl16_conv2d_stat_bytes.restype = sz
    lib.tcvn_sdxl16_conv2d_c64.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, vp]
    lib.tcvn_sdxl16_gn_stats.argtypes = [vp, i32, i32, i32, i32, f32, vp, vp, sz, vp]
    lib.tcvn_sdxl16_conv.argtypes = [vp, i64, i32, i32, vp, vp, i32, vp, i32, vp, vp, vp, i32, i32, i32, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("tcvn_abi_version",):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise TcvnError(f"{what} failed ({rc}): {load().tcvn_last_error().decode()}")


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise TcvnError(f"{what}: tensor is on {t.device}; this path has CUDA kernels only (no CPU fallback)")
