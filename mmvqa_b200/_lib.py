"""ctypes binding of libmmvqa_sm100.so (include/mmvqa.h).

The library is the only compute backend: there is no CPU or eager-PyTorch fallback.  If the
shared object is missing, ``lib()`` raises with the build command instead of degrading.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmmvqa_sm100.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_SERF, ACT_GELU, ACT_RELU = 0, 1, 2, 3
EPI_STORE, EPI_ACT, EPI_RESIDUAL, EPI_DACT, EPI_ACT_ROWSUM, EPI_DACT_SCALE = range(6)
ABI_VERSION = 4

vp, i64, i32, f32, u64 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint64


class GemmArgs(C.Structure):
    _fields_ = [
        ("dtype", i32), ("M", i32), ("N", i32), ("K", i32),
        ("A", vp), ("lda", i64), ("a_trans", i32),
        ("B", vp), ("ldb", i64), ("b_trans", i32),
        ("C", vp), ("ldc", i64), ("c_dtype", i32),
        ("bias", vp),
        ("epilogue", i32), ("act", i32),
        ("aux_in", vp), ("ld_aux_in", i64),
        ("aux_out", vp), ("ld_aux_out", i64),
        ("rowsum_out", vp), ("colsum_out", vp), ("rowscale", vp), ("scale", f32),
        ("accumulate", i32), ("split_k", i32),
        ("batch", i32), ("a_batch_rows", i64), ("b_batch_rows", i64), ("c_batch_stride", i64),
        ("dropout_p", f32), ("dropout_seed", u64),
        ("c_split_stride", i64), ("b_static", i32), ("trace", vp),
    ]


PREFETCH_MAX = 16


class PrefetchList(C.Structure):
    _fields_ = [("ptr", vp * PREFETCH_MAX), ("bytes", i64 * PREFETCH_MAX), ("n", i32)]


CAST_MULTI_MAX = 8


class CastList(C.Structure):
    _fields_ = [("src", vp * CAST_MULTI_MAX), ("dst", vp * CAST_MULTI_MAX), ("ld_src", i64 * CAST_MULTI_MAX),
                ("ld_dst", i64 * CAST_MULTI_MAX), ("rows", i64 * CAST_MULTI_MAX), ("cols", i32 * CAST_MULTI_MAX), ("src_bf16", i32 * CAST_MULTI_MAX),
                ("n", i32)]


class RfEncoderArgs(C.Structure):
    _fields_ = [
        ("B", i32), ("T", i32), ("hidden", i32), ("heads", i32), ("ff", i32), ("n_layers", i32),
        ("wkqv", C.POINTER(vp)), ("wproj", C.POINTER(vp)), ("w1", C.POINTER(vp)), ("w2", C.POINTER(vp)),
        ("b1", C.POINTER(vp)), ("b2", C.POINTER(vp)),
        ("ln1_w", C.POINTER(vp)), ("ln1_b", C.POINTER(vp)), ("ln2_w", C.POINTER(vp)), ("ln2_b", C.POINTER(vp)),
        ("x0", vp), ("xout", vp), ("kqv", vp), ("scores", vp), ("att", vp), ("y1", vp), ("x1", vp),
        ("hpre", vp), ("hact", vp), ("y2", vp),
        ("mean1", vp), ("rstd1", vp), ("mean2", vp), ("rstd2", vp),
        ("prev", vp), ("mask", vp),
        ("dropout_p1", f32), ("dropout_p2", f32), ("eps", f32), ("dropout_seed", u64), ("trace", vp),
    ]


class RfAttnBlockBwdArgs(C.Structure):
    _fields_ = [
        ("B", i32), ("T", i32), ("hidden", i32), ("heads", i32),
        ("dy_parts", vp), ("nparts", i32), ("part_stride", i64),
        ("dy_res", vp),
        ("y1", vp), ("mean1", vp), ("rstd1", vp), ("ln1_w", vp),
        ("wproj", vp), ("wkqv", vp),
        ("kqv", vp), ("scores", vp), ("dscores_in", vp),
        ("dpr", vp), ("dkqv", vp), ("dprev", vp), ("dxin", vp), ("dln1_w", vp), ("dln1_b", vp),
        ("dropout_p", f32), ("dropout_seed", u64),
    ]


class AdamDesc(C.Structure):
    _fields_ = [("p", vp), ("m", vp), ("v", vp), ("g", vp), ("bf16_out", vp), ("n", i64), ("flags", i64),
                ("row_live", vp), ("row_len", i64)]


# name -> (restype, argtypes); every symbol declared in include/mmvqa.h
SIGNATURES = {
    "mmvqa_abi_version": (i32, []),
    "mmvqa_last_error": (C.c_char_p, []),
    "mmvqa_device_sm": (i32, []),
    "mmvqa_launch_count": (i64, []),
    "mmvqa_set_dropout_counter": (i32, [vp]),
    "mmvqa_gemm": (i32, [C.POINTER(GemmArgs), vp]),
    "mmvqa_bias_act_fwd": (i32, [vp, vp, vp, i64, i32, i32, i32, vp]),
    "mmvqa_bias_act_bwd": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, vp]),
    "mmvqa_colsum": (i32, [vp, i64, vp, i64, i32, i32, vp]),
    "mmvqa_cast_pad_multi": (i32, [C.POINTER(CastList), vp]),
    "mmvqa_layernorm_bwd_partial_rows": (i32, [i64, i32, i32]),
    "mmvqa_layernorm_bwd_deferred": (i32, [vp, vp, i32, i64, vp, vp, vp, vp, vp, vp, f32, u64, i64, i32, i32, vp, i32, vp]),
    "mmvqa_ln_partials_reduce": (i32, [vp, i32, i32, vp, vp, vp, vp]),
    "mmvqa_vistok_pgrad_supported": (i32, [i32, i32, i32]),
    "mmvqa_vistok_fwd_pgrad": (i32, [vp, i64, vp, i64, vp, vp, i32, i32, i32, i32, i32, vp]),
    "mmvqa_vistok_dw": (i32, [vp, vp, f32, vp, i32, i32, i32, vp]),
    "mmvqa_l2_prefetch": (i32, [C.POINTER(PrefetchList), i32, vp]),
    "mmvqa_cast": (i32, [vp, i32, vp, i32, i64, vp]),
    "mmvqa_cast_pad": (i32, [vp, i32, i64, vp, i32, i64, i64, i32, vp]),
    "mmvqa_scale_by_device_scalar": (i32, [vp, i32, vp, f32, i64, vp]),
    "mmvqa_dropout": (i32, [vp, vp, i64, f32, u64, i32, vp]),
    "mmvqa_add_layernorm_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, f32, i32, vp]),
    "mmvqa_layernorm_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, u64, i64, i32, i32, vp]),
    "mmvqa_add_layernorm_fwd_parts": (i32, [vp, i32, i64, vp, vp, vp, vp, vp, vp, vp, i64, i32, f32, f32, u64, i32, vp]),
    "mmvqa_layernorm_bwd_parts": (i32, [vp, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, u64, i64, i32, i32, vp]),
    "mmvqa_mhsa_fwd": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, f32, u64, i32, vp]),
    "mmvqa_mhsa_bwd": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, f32, u64, i32, vp]),
    "mmvqa_rf_attn_fwd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "mmvqa_rf_attn_fwd_fused": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "mmvqa_rf_attn_bwd_fused": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "mmvqa_rf_attn_bwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "mmvqa_rf_attn_block_bwd_supported": (i32, [i32, i32, i32, i32]),
    "mmvqa_rf_attn_block_bwd": (i32, [C.POINTER(RfAttnBlockBwdArgs), vp]),
    "mmvqa_rf_encoder_fwd_supported": (i32, [i32, i32, i32, i32, i32, i32]),
    "mmvqa_rf_encoder_fwd": (i32, [C.POINTER(RfEncoderArgs), vp]),
    "mmvqa_embed_ln_scatter_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32,
                                         f32, u64, i32, vp]),
    "mmvqa_embed_ln_scatter_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32,
                                         i32, i32, f32, u64, i32, vp]),
    "mmvqa_masked_mean_fwd": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "mmvqa_masked_mean_bwd": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "mmvqa_l2norm_fwd": (i32, [vp, vp, vp, i32, i32, vp]),
    "mmvqa_l2norm_bwd": (i32, [vp, vp, vp, vp, i32, i32, vp]),
    "mmvqa_asl_fwd_bwd": (i32, [vp, i64, vp, vp, vp, vp, i32, i32, f32, f32, f32, i32, vp]),
    "mmvqa_ce_fwd_bwd": (i32, [vp, i64, vp, vp, vp, i64, i64, i32, f32, i32, vp]),
    "mmvqa_supcon_rows": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, f32, f32, vp]),
    "mmvqa_ce_chunk_stats": (i32, [vp, i64, vp, i64, i32, i32, vp, vp, vp, i32, vp]),
    "mmvqa_ce_chunk_grad": (i32, [vp, i64, vp, i64, i32, i32, vp, vp, vp, vp, i64, i32, vp]),
    "mmvqa_jaccard_mask": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "mmvqa_adam_step": (i32, [C.POINTER(AdamDesc), i32, f32, f32, f32, f32, f32, i32, vp, f32, i32, i32, vp]),
    "mmvqa_multimem_allreduce": (i32, [vp, vp, i32, i32, i64, i32, i32, vp]),
    "mmvqa_mark_rows": (i32, [vp, vp, i64, i64, vp]),
    "mmvqa_adam_step_dev": (i32, [C.POINTER(AdamDesc), i32, vp, f32, f32, f32, f32, vp, i32, i32, vp]),
}

_LIB = None


class MMVQAError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the shared library; fail loudly if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise MMVQAError(
                f"{LIB_PATH} is missing: build it with `python -m mmvqa_b200.build` "
                "(nvcc, sm_100a).  There is no CPU / eager fallback for this path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        got = handle.mmvqa_abi_version()
        if got != ABI_VERSION:
            raise MMVQAError(f"libmmvqa_sm100.so ABI {got} != expected {ABI_VERSION}; rebuild")
        _LIB = handle
    return _LIB


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().mmvqa_last_error()
        raise MMVQAError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(lib().mmvqa_launch_count())


def set_dropout_counter(counter) -> None:
    """register (tensor) / clear (None) the device uint64 step counter mixed into every dropout seed of the kernels
    launched from now on (CUDA-graph replay: mmvqa_b200.graph.GraphedTrainStep)."""
    check(lib().mmvqa_set_dropout_counter(None if counter is None else counter.data_ptr()), "set_dropout_counter")
