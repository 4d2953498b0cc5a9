"""Thin tensor-level wrappers over the C ABI (one Python function per entry point of include/mmvqa.h).

Every function takes CUDA tensors, passes raw device pointers + the current CUDA stream, and raises
``MMVQAError`` on a non-zero return code.  Nothing here allocates behind the caller's back except
where the docstring says so, and nothing falls back to torch math.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from ._lib import (ACT_GELU, ACT_NONE, ACT_RELU, ACT_SERF, BF16, EPI_ACT, EPI_ACT_ROWSUM, EPI_DACT, EPI_DACT_SCALE,
                   EPI_RESIDUAL, EPI_STORE, F32)

Tensor = torch.Tensor
ACT_CODES = {"none": ACT_NONE, "serf": ACT_SERF, "gelu": ACT_GELU, "relu": ACT_RELU}
_GEMM_RECORD = None     # bench.py instrumentation: list of (signature, flops, GemmArgs, keep-alive tensors)


def gemm_record(enable: bool):
    """Start / stop recording every mmvqa_gemm launch (its argument struct and the tensors it points to) so that
    bench.py can replay each distinct problem in isolation and time it (roofline measurement).  Returns the list
    when stopping."""
    global _GEMM_RECORD
    out = _GEMM_RECORD
    _GEMM_RECORD = [] if enable else None
    return out


def gemm_replay(args) -> None:
    L.check(L.lib().mmvqa_gemm(C.byref(args), _stream()), "mmvqa_gemm")


def dtype_code(t) -> int:
    dt = t.dtype if isinstance(t, torch.Tensor) else t
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise L.MMVQAError(f"unsupported dtype {dt} (float32 or bfloat16 only)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise L.MMVQAError("mmvqa_b200 kernels need CUDA tensors: there is no CPU fallback for this path")
    return t.data_ptr()


def _cont(t: Tensor, name: str) -> Tensor:
    if not t.is_contiguous():
        raise L.MMVQAError(f"{name} must be contiguous")
    return t


# ------------------------------------------------------------------------------- GEMM
def gemm(M: int, N: int, K: int, A: Tensor, lda: int, a_trans: bool, B: Tensor, ldb: int, b_trans: bool,
         Cout: Optional[Tensor], ldc: int, *, bias: Optional[Tensor] = None, epilogue: int = EPI_STORE,
         act: int = ACT_NONE, aux_in: Optional[Tensor] = None, ld_aux_in: int = 0, aux_out: Optional[Tensor] = None,
         ld_aux_out: int = 0, rowsum_out: Optional[Tensor] = None, colsum_out: Optional[Tensor] = None,
         rowscale: Optional[Tensor] = None,
         scale: float = 1.0, accumulate: bool = False, split_k: int = 1, batch: int = 1, a_batch_rows: int = 0,
         b_batch_rows: int = 0, c_batch_stride: int = 0, dropout_p: float = 0.0, dropout_seed: int = 0,
         c_split_stride: int = 0, b_static: bool = False, trace: Optional[Tensor] = None) -> None:
    """C[M,N] = epilogue(sum_k opA(A)[m,k] opB(B)[n,k]); see include/mmvqa.h for the epilogues."""
    if A.dtype != B.dtype:
        raise L.MMVQAError(f"gemm operands differ in dtype: {A.dtype} vs {B.dtype}")
    if bias is not None and bias.dtype != torch.float32:
        raise L.MMVQAError("gemm bias must be float32")
    a = L.GemmArgs()
    a.dtype = dtype_code(A)
    a.M, a.N, a.K = M, N, K
    a.A, a.lda, a.a_trans = _p(A), lda, int(a_trans)
    a.B, a.ldb, a.b_trans = _p(B), ldb, int(b_trans)
    a.C, a.ldc = _p(Cout), ldc
    a.c_dtype = dtype_code(Cout) if Cout is not None else a.dtype
    a.bias = _p(bias)
    a.epilogue, a.act = epilogue, act
    a.aux_in, a.ld_aux_in = _p(aux_in), ld_aux_in
    a.aux_out, a.ld_aux_out = _p(aux_out), ld_aux_out
    a.rowsum_out, a.colsum_out, a.rowscale, a.scale = _p(rowsum_out), _p(colsum_out), _p(rowscale), scale
    a.accumulate, a.split_k = int(accumulate), split_k
    a.batch, a.a_batch_rows, a.b_batch_rows, a.c_batch_stride = batch, a_batch_rows, b_batch_rows, c_batch_stride
    a.dropout_p, a.dropout_seed = dropout_p, dropout_seed
    a.c_split_stride = c_split_stride
    a.b_static = int(b_static)
    a.trace = _p(trace)
    for t in (aux_in, aux_out):
        if t is not None and t.dtype != A.dtype:
            raise L.MMVQAError("gemm aux tensors must have the operand dtype")
    if _GEMM_RECORD is not None:
        sig = (a.dtype, M, N, K, int(a_trans), int(b_trans), epilogue, act, batch, split_k, int(accumulate), a.c_dtype, c_split_stride,
               bias is not None, aux_out is not None, colsum_out is not None, rowscale is not None, dropout_p > 0)
        _GEMM_RECORD.append((sig, 2.0 * M * N * K * max(batch, 1), a,
                             (A, B, Cout, bias, aux_in, aux_out, rowsum_out, colsum_out, rowscale)))
    L.check(L.lib().mmvqa_gemm(C.byref(a), _stream()), "mmvqa_gemm")


# ------------------------------------------------------------------------------- elementwise
def bias_act_fwd(x: Tensor, bias: Optional[Tensor], act: int, out: Optional[Tensor] = None) -> Tensor:
    _cont(x, "x")
    cols = x.shape[-1]
    rows = x.numel() // cols if cols else 0
    y = torch.empty_like(x) if out is None else out
    L.check(L.lib().mmvqa_bias_act_fwd(_p(x), _p(bias), _p(y), rows, cols, act, dtype_code(x), _stream()), "bias_act_fwd")
    return y


def bias_act_bwd(x: Tensor, bias: Optional[Tensor], dy: Tensor, act: int) -> Tensor:
    _cont(x, "x"), _cont(dy, "dy")
    cols = x.shape[-1]
    rows = x.numel() // cols if cols else 0
    dx = torch.empty_like(x)
    L.check(L.lib().mmvqa_bias_act_bwd(_p(x), _p(bias), _p(dy), _p(dx), rows, cols, act, dtype_code(x), _stream()),
            "bias_act_bwd")
    return dx


def colsum(x: Tensor, rows: int, cols: int, ldx: Optional[int] = None) -> Tensor:
    out = torch.empty(cols, device=x.device, dtype=torch.float32)
    L.check(L.lib().mmvqa_colsum(_p(x), cols if ldx is None else ldx, _p(out), rows, cols, dtype_code(x), _stream()), "colsum")
    return out


def vistok_pgrad_supported(M: int, HW: int, C: int) -> bool:
    return bool(L.lib().mmvqa_vistok_pgrad_supported(M, HW, C))


def vistok_fwd_pgrad(w: Tensor, f2d: Tensor, ldf: int, vis: Tensor, B: int, M: int, HW: int, C: int, act: int) -> Tensor:
    """vis [B, M] += mean_hw act(W f);  returns P [B, M, C] fp32 = sum_hw act'(W f) f  (see include/mmvqa.h)."""
    if w.dtype != torch.bfloat16 or f2d.dtype != torch.bfloat16:
        raise L.MMVQAError("vistok_fwd_pgrad: bf16 operands only")
    _cont(w, "w")
    pg = torch.zeros(B, M, C, device=w.device, dtype=torch.float32)
    L.check(L.lib().mmvqa_vistok_fwd_pgrad(_p(w), C, _p(f2d), ldf, _p(vis), _p(pg), M, HW, C, B, act, _stream()),
            "vistok_fwd_pgrad")
    return pg


def vistok_dw(pgrad: Tensor, dv: Tensor, scale: float) -> Tensor:
    """dW [M, C] = scale * sum_b dv[b, m] P[b, m, c]"""
    B, M, C = pgrad.shape
    _cont(dv, "dv")
    dw = torch.empty(M, C, device=pgrad.device, dtype=torch.float32)
    L.check(L.lib().mmvqa_vistok_dw(_p(pgrad), _p(dv), scale, _p(dw), B, M, C, _stream()), "vistok_dw")
    return dw


def l2_prefetch(tensors, ctas: int = 16) -> None:
    """L2 prefetch hints for the storage of `tensors` (contiguous ranges) on the current stream; no result."""
    ts = [t for t in tensors if t is not None and t.numel() > 0]
    for i in range(0, len(ts), L.PREFETCH_MAX):
        pl = L.PrefetchList()
        chunk = ts[i:i + L.PREFETCH_MAX]
        for j, t in enumerate(chunk):
            pl.ptr[j] = _p(t)
            pl.bytes[j] = t.numel() * t.element_size()
        pl.n = len(chunk)
        L.check(L.lib().mmvqa_l2_prefetch(C.byref(pl), ctas, _stream()), "l2_prefetch")


def cast(src: Tensor, dst_dtype: torch.dtype, out: Optional[Tensor] = None) -> Tensor:
    _cont(src, "src")
    dst = torch.empty(src.shape, device=src.device, dtype=dst_dtype) if out is None else out
    L.check(L.lib().mmvqa_cast(_p(src), dtype_code(src), _p(dst), dtype_code(dst), src.numel(), _stream()), "cast")
    return dst


def cast_pad(src: Tensor, rows: int, cols: int, ld_src: int, dst_dtype: torch.dtype, ld_dst: int) -> Tensor:
    dst = torch.empty(rows, ld_dst, device=src.device, dtype=dst_dtype)
    L.check(L.lib().mmvqa_cast_pad(_p(src), dtype_code(src), ld_src, _p(dst), dtype_code(dst), ld_dst, rows, cols, _stream()),
            "cast_pad")
    return dst


def cast_pad_multi(srcs, ld_dsts) -> list:
    """fp32 / bf16 [rows, cols] contiguous matrices -> bf16 [rows, ld_dst] (columns cols..ld_dst zero) in ONE launch."""
    outs = []
    for i in range(0, len(srcs), L.CAST_MULTI_MAX):
        cl = L.CastList()
        chunk = srcs[i:i + L.CAST_MULTI_MAX]
        for j, (t, ld) in enumerate(zip(chunk, ld_dsts[i:i + L.CAST_MULTI_MAX])):
            if t.dtype not in (torch.float32, torch.bfloat16) or t.dim() != 2 or not t.is_contiguous():
                raise L.MMVQAError("cast_pad_multi: contiguous 2-D float32 / bfloat16 sources only")
            rows, cols = t.shape
            dst = torch.empty(rows, ld, device=t.device, dtype=torch.bfloat16)
            cl.src[j], cl.dst[j], cl.ld_src[j], cl.ld_dst[j], cl.rows[j], cl.cols[j] = _p(t), _p(dst), cols, ld, rows, cols
            cl.src_bf16[j] = 1 if t.dtype == torch.bfloat16 else 0
            outs.append(dst)
        cl.n = len(chunk)
        L.check(L.lib().mmvqa_cast_pad_multi(C.byref(cl), _stream()), "cast_pad_multi")
    return outs


def scale_(x: Tensor, scalar: Optional[Tensor], host_factor: float = 1.0) -> Tensor:
    _cont(x, "x")
    if scalar is not None and (scalar.dtype != torch.float32 or scalar.numel() != 1):
        raise L.MMVQAError("scale_: scalar must be a 1-element float32 tensor")
    L.check(L.lib().mmvqa_scale_by_device_scalar(_p(x), dtype_code(x), _p(scalar), host_factor, x.numel(), _stream()), "scale")
    return x


def dropout(x: Tensor, p: float, seed: int) -> Tensor:
    _cont(x, "x")
    y = torch.empty_like(x)
    L.check(L.lib().mmvqa_dropout(_p(x), _p(y), x.numel(), p, seed, dtype_code(x), _stream()), "dropout")
    return y


def add_layernorm_fwd(x: Tensor, res: Optional[Tensor], gamma: Tensor, beta: Tensor, eps: float, want_sum: bool):
    """returns (y, xsum or None, mean, rstd)."""
    _cont(x, "x")
    cols = x.shape[-1]
    rows = x.numel() // cols
    y = torch.empty_like(x)
    xsum = torch.empty_like(x) if want_sum else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    L.check(L.lib().mmvqa_add_layernorm_fwd(_p(x), _p(res), _p(gamma), _p(beta), _p(y), _p(xsum), _p(mean), _p(rstd), rows,
                                           cols, eps, dtype_code(x), _stream()), "add_layernorm_fwd")
    return y, xsum, mean, rstd


def layernorm_bwd(dy: Tensor, xsum: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor, dx_extra: Optional[Tensor],
                  dgamma: Tensor, dbeta: Tensor, *, want_drop: bool = False, dxsum: Optional[Tensor] = None,
                  dropout_p: float = 0.0, dropout_seed: int = 0):
    """dgamma / dbeta (and dxsum) are ACCUMULATED into (caller zero-fills).  Returns dx, or (dx, dropout(dx)) when
    want_drop."""
    _cont(dy, "dy"), _cont(xsum, "xsum")
    cols = dy.shape[-1]
    rows = dy.numel() // cols
    dx = torch.empty_like(dy)
    dxd = torch.empty_like(dy) if want_drop else None
    L.check(L.lib().mmvqa_layernorm_bwd(_p(dy), _p(xsum), _p(gamma), _p(mean), _p(rstd), _p(dx_extra), _p(dx), _p(dgamma),
                                       _p(dbeta), _p(dxd), _p(dxsum), dropout_p, dropout_seed, rows, cols, dtype_code(dy),
                                       _stream()), "layernorm_bwd")
    return (dx, dxd) if want_drop else dx


def add_layernorm_fwd_parts(parts: Tensor, res: Optional[Tensor], gamma: Tensor, beta: Tensor, eps: float,
                            out_dtype: torch.dtype, dropout_p: float = 0.0, dropout_seed: int = 0):
    """parts [nparts, rows, cols] fp32 split-K slabs -> (y, s, mean, rstd) with s = dropout(sum parts) + res."""
    _cont(parts, "parts")
    nparts, rows, cols = parts.shape
    if parts.dtype != torch.float32:
        raise L.MMVQAError("split-K partials must be float32")
    y = torch.empty(rows, cols, device=parts.device, dtype=out_dtype)
    s = torch.empty_like(y)
    mean = torch.empty(rows, device=parts.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=parts.device, dtype=torch.float32)
    L.check(L.lib().mmvqa_add_layernorm_fwd_parts(_p(parts), nparts, rows * cols, _p(res), _p(gamma), _p(beta), _p(y), _p(s),
                                                 _p(mean), _p(rstd), rows, cols, eps, dropout_p, dropout_seed,
                                                 dtype_code(out_dtype), _stream()), "add_layernorm_fwd_parts")
    return y, s, mean, rstd


def layernorm_bwd_parts(dy_parts: Tensor, dy_res: Optional[Tensor], xsum: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor,
                        dgamma: Tensor, dbeta: Tensor, *, want_drop: bool = False, dxsum: Optional[Tensor] = None,
                        dropout_p: float = 0.0, dropout_seed: int = 0):
    """layernorm_bwd whose incoming gradient is sum(dy_parts [nparts, rows, cols] fp32) + dy_res."""
    _cont(dy_parts, "dy_parts"), _cont(xsum, "xsum")
    nparts, rows, cols = dy_parts.shape
    dx = torch.empty_like(xsum)
    dxd = torch.empty_like(xsum) if want_drop else None
    L.check(L.lib().mmvqa_layernorm_bwd_parts(_p(dy_parts), nparts, rows * cols, _p(dy_res), _p(xsum), _p(gamma), _p(mean),
                                             _p(rstd), _p(dx), _p(dgamma), _p(dbeta), _p(dxd), _p(dxsum), dropout_p,
                                             dropout_seed, rows, cols, dtype_code(xsum), _stream()), "layernorm_bwd_parts")
    return (dx, dxd) if want_drop else dx


def layernorm_bwd_deferred(dy: Optional[Tensor], dy_parts: Optional[Tensor], xsum: Tensor, gamma: Tensor, mean: Tensor,
                           rstd: Tensor, *, want_drop: bool = False, dropout_p: float = 0.0, dropout_seed: int = 0):
    """LayerNorm backward whose gamma / beta / bias column sums are left as per-CTA partials.  Returns
    (dx, dx_drop or None, partials [n, 3, cols]) or None when this shape has no deferred form (use layernorm_bwd).
    The incoming gradient is dy (+ sum of the fp32 slabs dy_parts [nparts, rows, cols])."""
    cols = xsum.shape[-1]
    rows = xsum.numel() // cols
    n = L.lib().mmvqa_layernorm_bwd_partial_rows(rows, cols, dtype_code(xsum))
    if n <= 0:
        return None
    _cont(xsum, "xsum")
    nparts, stride = 0, 0
    if dy_parts is not None:
        _cont(dy_parts, "dy_parts")
        nparts, stride = dy_parts.shape[0], rows * cols
    if dy is not None:
        _cont(dy, "dy")
    dx = torch.empty_like(xsum)
    dxd = torch.empty_like(xsum) if want_drop else None
    partials = torch.empty(n, 3, cols, device=xsum.device, dtype=torch.float32)
    L.check(L.lib().mmvqa_layernorm_bwd_deferred(_p(dy), _p(dy_parts), nparts, stride, _p(xsum), _p(gamma), _p(mean), _p(rstd),
                                                _p(dx), _p(dxd), dropout_p, dropout_seed, rows, cols, dtype_code(xsum),
                                                _p(partials), n, _stream()), "layernorm_bwd_deferred")
    return dx, dxd, partials


def ln_partials_reduce(partials: Tensor, dgamma: Optional[Tensor], dbeta: Optional[Tensor], dxsum: Optional[Tensor]) -> None:
    """dgamma / dbeta / dxsum += column sums held in `partials` (layernorm_bwd_deferred)."""
    n, _, cols = partials.shape
    L.check(L.lib().mmvqa_ln_partials_reduce(_p(partials), n, cols, _p(dgamma), _p(dbeta), _p(dxsum), _stream()),
            "ln_partials_reduce")


# ------------------------------------------------------------------------------- attention
def rf_attn_fwd(kqv: Tensor, prev: Optional[Tensor], mask: Optional[Tensor], B: int, T: int, heads: int, d: int):
    out = torch.empty(B * T, heads * d, device=kqv.device, dtype=kqv.dtype)
    scores = torch.empty(B, heads, T, T, device=kqv.device, dtype=torch.float32)
    L.check(L.lib().mmvqa_rf_attn_fwd(_p(kqv), _p(prev), _p(mask), _p(out), _p(scores), B, T, heads, d, dtype_code(kqv),
                                     _stream()), "rf_attn_fwd")
    return out, scores


def rf_attn_fwd_fused(x: Tensor, wkqv: Tensor, prev: Optional[Tensor], mask: Optional[Tensor], B: int, T: int, heads: int,
                      d: int):
    """kqv projection + residual attention in one launch (bf16): returns (out, scores, kqv)."""
    out = torch.empty(B * T, heads * d, device=x.device, dtype=x.dtype)
    scores = torch.empty(B, heads, T, T, device=x.device, dtype=torch.float32)
    kqv = torch.empty(B * T * heads, 3 * d, device=x.device, dtype=x.dtype)
    L.check(L.lib().mmvqa_rf_attn_fwd_fused(_p(x), _p(wkqv), _p(prev), _p(mask), _p(out), _p(scores), _p(kqv), B, T, heads, d,
                                           dtype_code(x), _stream()), "rf_attn_fwd_fused")
    return out, scores, kqv


RF_TRACE_BUFFER: Optional[Tensor] = None      # int64 [4, 16, 16] CUDA tensor: phase stamps of CTA 0 (tools/rf_encoder_check.py)


def rf_encoder_supported(B: int, T: int, hidden: int, heads: int, ff: int, n_layers: int) -> bool:
    return bool(L.lib().mmvqa_rf_encoder_fwd_supported(B, T, hidden, heads, ff, n_layers))


def rf_encoder_fwd(x0: Tensor, layers, prev: Optional[Tensor], mask: Optional[Tensor], B: int, T: int, heads: int, p1: float,
                   p2: float, eps: float, seed: int):
    """whole RealFormer encoder forward in one launch (bf16).  `layers` = per layer (wkqv, wproj, w1, w2 [bf16 operand
    copies], b1, b2, ln1_w, ln1_b, ln2_w, ln2_b [fp32]).  Returns the dict of stacked [n_layers, ...] buffers the backward
    pass reads (same layouts as the per-operator path)."""
    n = len(layers)
    M, H = x0.shape
    F4 = layers[0][2].shape[0]
    d = H // heads
    dev, bf = x0.device, torch.bfloat16
    out = dict(
        xout=torch.empty(n, M, H, device=dev, dtype=bf), kqv=torch.empty(n, M * heads, 3 * d, device=dev, dtype=bf),
        scores=torch.empty(n, B, heads, T, T, device=dev, dtype=torch.float32),
        att=torch.empty(n, M, H, device=dev, dtype=bf), y1=torch.empty(n, M, H, device=dev, dtype=bf),
        x1=torch.empty(n, M, H, device=dev, dtype=bf), hpre=torch.empty(n, M, F4, device=dev, dtype=bf),
        hact=torch.empty(n, M, F4, device=dev, dtype=bf), y2=torch.empty(n, M, H, device=dev, dtype=bf),
        mean1=torch.empty(n, M, device=dev, dtype=torch.float32), rstd1=torch.empty(n, M, device=dev, dtype=torch.float32),
        mean2=torch.empty(n, M, device=dev, dtype=torch.float32), rstd2=torch.empty(n, M, device=dev, dtype=torch.float32))
    a = L.RfEncoderArgs()
    a.B, a.T, a.hidden, a.heads, a.ff, a.n_layers = B, T, H, heads, F4, n
    arrs = []
    for k, name in enumerate(("wkqv", "wproj", "w1", "w2", "b1", "b2", "ln1_w", "ln1_b", "ln2_w", "ln2_b")):
        for lay in layers:
            t = lay[k]
            want = bf if k < 4 else torch.float32
            if t.dtype != want or not t.is_contiguous() or not t.is_cuda:
                raise L.MMVQAError("rf_encoder_fwd: parameter %s must be a contiguous CUDA %s tensor" % (name, want))
        arr = (L.vp * n)(*[lay[k].data_ptr() for lay in layers])
        arrs.append(arr)
        setattr(a, name, C.cast(arr, C.POINTER(L.vp)))
    a.x0 = _p(_cont(x0, "x0"))
    for name in ("xout", "kqv", "scores", "att", "y1", "x1", "hpre", "hact", "y2", "mean1", "rstd1", "mean2", "rstd2"):
        setattr(a, name, out[name].data_ptr())
    a.prev = _p(prev)
    a.mask = _p(mask)
    a.dropout_p1, a.dropout_p2, a.eps, a.dropout_seed = p1, p2, eps, seed & 0xFFFFFFFFFFFFFFFF
    a.trace = _p(RF_TRACE_BUFFER)
    L.check(L.lib().mmvqa_rf_encoder_fwd(C.byref(a), _stream()), "rf_encoder_fwd")
    return out


def rf_attn_block_bwd_supported(B: int, T: int, hidden: int, heads: int) -> bool:
    return bool(L.lib().mmvqa_rf_attn_block_bwd_supported(B, T, hidden, heads))


def rf_attn_block_bwd(dy_parts: Tensor, dy_res: Optional[Tensor], y1: Tensor, mean1: Tensor, rstd1: Tensor, ln1_w: Tensor,
                      wproj: Tensor, wkqv: Tensor, kqv: Tensor, scores: Tensor, dscores_in: Optional[Tensor], want_dprev: bool,
                      dln1_w: Tensor, dln1_b: Tensor, B: int, T: int, heads: int, p: float, seed: int):
    """LN1 backward + proj dgrad + residual-attention backward + kqv dgrad of one layer in one cluster launch (bf16).
    dy_parts [nparts, M, H] fp32.  Returns (dpr, dkqv, dprev, dxin); dln1_w / dln1_b are accumulated in place."""
    nparts, M, H = dy_parts.shape
    dpr = torch.empty(M, H, device=y1.device, dtype=y1.dtype)
    dkqv = torch.empty_like(kqv)
    dxin = torch.empty(M, H, device=y1.device, dtype=y1.dtype)
    dprev = torch.empty_like(scores) if want_dprev else None
    a = L.RfAttnBlockBwdArgs()
    a.B, a.T, a.hidden, a.heads = B, T, H, heads
    a.dy_parts, a.nparts, a.part_stride = _p(_cont(dy_parts, "dy_parts")), nparts, M * H
    a.dy_res = _p(dy_res)
    a.y1, a.mean1, a.rstd1, a.ln1_w = _p(_cont(y1, "y1")), _p(mean1), _p(rstd1), _p(ln1_w)
    a.wproj, a.wkqv = _p(_cont(wproj, "wproj")), _p(_cont(wkqv, "wkqv"))
    a.kqv, a.scores, a.dscores_in = _p(_cont(kqv, "kqv")), _p(_cont(scores, "scores")), _p(dscores_in)
    a.dpr, a.dkqv, a.dprev, a.dxin = _p(dpr), _p(dkqv), _p(dprev), _p(dxin)
    a.dln1_w, a.dln1_b = _p(dln1_w), _p(dln1_b)
    a.dropout_p, a.dropout_seed = p, seed & 0xFFFFFFFFFFFFFFFF
    L.check(L.lib().mmvqa_rf_attn_block_bwd(C.byref(a), _stream()), "rf_attn_block_bwd")
    return dpr, dkqv, dprev, dxin


def rf_attn_bwd(kqv: Tensor, scores: Tensor, dout: Tensor, dscores_in: Optional[Tensor], want_dprev: bool, B: int, T: int,
                heads: int, d: int):
    dkqv = torch.empty_like(kqv)
    dprev = torch.empty_like(scores) if want_dprev else None
    L.check(L.lib().mmvqa_rf_attn_bwd(_p(kqv), _p(scores), _p(dout), _p(dscores_in), _p(dkqv), _p(dprev), B, T, heads, d,
                                     dtype_code(kqv), _stream()), "rf_attn_bwd")
    return dkqv, dprev


def rf_attn_bwd_fused(kqv: Tensor, scores: Tensor, dout: Tensor, dscores_in: Optional[Tensor], want_dprev: bool,
                      wkqv: Tensor, dres: Optional[Tensor], B: int, T: int, heads: int, d: int):
    """attention backward + input gradient of the kqv projection (bf16): returns (dkqv, dprev, dx) with
    dx [B*T, heads*d] = dkqv . wkqv + dres."""
    dkqv = torch.empty_like(kqv)
    dprev = torch.empty_like(scores) if want_dprev else None
    dx = torch.empty(B * T, heads * d, device=kqv.device, dtype=kqv.dtype)
    L.check(L.lib().mmvqa_rf_attn_bwd_fused(_p(kqv), _p(scores), _p(dout), _p(dscores_in), _p(dkqv), _p(dprev), _p(wkqv),
                                           _p(dres), _p(dx), B, T, heads, d, dtype_code(kqv), _stream()), "rf_attn_bwd_fused")
    return dkqv, dprev, dx


def mhsa_fwd(qkv: Tensor, mask: Optional[Tensor], B: int, T: int, heads: int, d: int, p: float, seed: int):
    out = torch.empty(B * T, heads * d, device=qkv.device, dtype=qkv.dtype)
    probs = torch.empty(B, heads, T, T, device=qkv.device, dtype=qkv.dtype)
    L.check(L.lib().mmvqa_mhsa_fwd(_p(qkv), _p(mask), _p(out), _p(probs), B, T, heads, d, p, seed, dtype_code(qkv), _stream()),
            "mhsa_fwd")
    return out, probs


def mhsa_bwd(qkv: Tensor, probs: Tensor, dout: Tensor, B: int, T: int, heads: int, d: int, p: float, seed: int) -> Tensor:
    dqkv = torch.empty_like(qkv)
    L.check(L.lib().mmvqa_mhsa_bwd(_p(qkv), _p(probs), _p(dout), _p(dqkv), B, T, heads, d, p, seed, dtype_code(qkv), _stream()),
            "mhsa_bwd")
    return dqkv


# ------------------------------------------------------------------------------- fusion / pooling
def embed_ln_scatter_fwd(ids, seg, word, pos, typ, gamma, beta, vis, out_dtype, eps, p, seed):
    B, T = ids.shape
    H = word.shape[1]
    nvis = 0 if vis is None else vis.shape[0]
    h = torch.empty(B, T, H, device=word.device, dtype=out_dtype)
    mean = torch.empty(B * T, device=word.device, dtype=torch.float32)
    rstd = torch.empty(B * T, device=word.device, dtype=torch.float32)
    L.check(L.lib().mmvqa_embed_ln_scatter_fwd(_p(ids), _p(seg), _p(word), _p(pos), _p(typ), _p(gamma), _p(beta), _p(vis), _p(h),
                                              _p(mean), _p(rstd), B, T, H, nvis, word.shape[0], eps, p, seed, dtype_code(h),
                                              _stream()), "embed_ln_scatter_fwd")
    return h, mean, rstd


def embed_ln_scatter_bwd(dh, ids, seg, word, pos, typ, gamma, mean, rstd, dword, dpos, dtyp, dgamma, dbeta, dvis, nvis,
                         padding_idx, p, seed):
    B, T = ids.shape
    H = word.shape[1]
    L.check(L.lib().mmvqa_embed_ln_scatter_bwd(_p(dh), _p(ids), _p(seg), _p(word), _p(pos), _p(typ), _p(gamma), _p(mean),
                                              _p(rstd), _p(dword), _p(dpos), _p(dtyp), _p(dgamma), _p(dbeta), _p(dvis), B, T, H,
                                              nvis, padding_idx, p, seed, dtype_code(dh), _stream()), "embed_ln_scatter_bwd")


def masked_mean_fwd(h: Tensor, mask: Tensor) -> Tensor:
    B, T, H = h.shape
    out = torch.empty(B, H, device=h.device, dtype=h.dtype)
    L.check(L.lib().mmvqa_masked_mean_fwd(_p(_cont(h, "h")), _p(mask), _p(out), B, T, H, dtype_code(h), _stream()), "masked_mean_fwd")
    return out


def masked_mean_bwd(dout: Tensor, mask: Tensor, T: int) -> Tensor:
    B, H = dout.shape
    dh = torch.empty(B, T, H, device=dout.device, dtype=dout.dtype)
    L.check(L.lib().mmvqa_masked_mean_bwd(_p(_cont(dout, "dout")), _p(mask), _p(dh), B, T, H, dtype_code(dout), _stream()),
            "masked_mean_bwd")
    return dh


def l2norm_fwd(x: Tensor):
    rows, cols = x.shape
    y = torch.empty_like(x)
    inv = torch.empty(rows, device=x.device, dtype=torch.float32)
    L.check(L.lib().mmvqa_l2norm_fwd(_p(_cont(x, "x")), _p(y), _p(inv), rows, cols, _stream()), "l2norm_fwd")
    return y, inv


def l2norm_bwd(y: Tensor, inv: Tensor, dy: Tensor) -> Tensor:
    rows, cols = y.shape
    dx = torch.empty_like(y)
    L.check(L.lib().mmvqa_l2norm_bwd(_p(y), _p(inv), _p(_cont(dy, "dy")), _p(dx), rows, cols, _stream()), "l2norm_bwd")
    return dx


# ------------------------------------------------------------------------------- losses
def asl_fwd_bwd(logits: Tensor, ld: int, target: Tensor, C_: int, gp: float, gn: float, eps: float, want_grad: bool,
                want_tc: bool):
    B = target.shape[0]
    loss_rows = torch.empty(B, device=logits.device, dtype=torch.float32)
    dl = torch.empty(B, C_, device=logits.device, dtype=torch.float32) if want_grad else None
    tc = torch.empty(B, C_, device=logits.device, dtype=torch.float32) if want_tc else None
    L.check(L.lib().mmvqa_asl_fwd_bwd(_p(logits), ld, _p(target), _p(loss_rows), _p(dl), _p(tc), B, C_, gp, gn, eps,
                                     dtype_code(logits), _stream()), "asl_fwd_bwd")
    return loss_rows, dl, tc


def ce_fwd_bwd(logits: Tensor, ld: int, target: Tensor, rows: int, C_: int, scale: float, dlogits: Optional[Tensor], ld_d: int):
    loss_rows = torch.empty(rows, device=logits.device, dtype=torch.float32)
    L.check(L.lib().mmvqa_ce_fwd_bwd(_p(logits), ld, _p(target), _p(loss_rows), _p(dlogits), ld_d, rows, C_, scale,
                                    dtype_code(logits), _stream()), "ce_fwd_bwd")
    return loss_rows


def ce_chunk_stats(chunk: Tensor, ld: int, target: Tensor, rows: int, col0: int, Vc: int, rowmax: Tensor, rowsum: Tensor,
                   tgt_logit: Tensor, first: bool) -> None:
    L.check(L.lib().mmvqa_ce_chunk_stats(_p(chunk), ld, _p(target), rows, col0, Vc, _p(rowmax), _p(rowsum), _p(tgt_logit),
                                        int(first), _stream()), "ce_chunk_stats")


def ce_chunk_grad(chunk: Tensor, ld: int, target: Tensor, rows: int, col0: int, Vc: int, rowmax: Tensor, rowsum: Tensor,
                  row_scale: Tensor, dl: Tensor, ld_d: int) -> None:
    L.check(L.lib().mmvqa_ce_chunk_grad(_p(chunk), ld, _p(target), rows, col0, Vc, _p(rowmax), _p(rowsum), _p(row_scale),
                                       _p(dl), ld_d, dtype_code(dl), _stream()), "ce_chunk_grad")


def supcon_rows(raw: Tensor, mask: Optional[Tensor], bsz: int, row_offset: int, temperature: float, base_temperature: float,
                want_grad: bool):
    R, N = raw.shape
    loss_rows = torch.empty(R, device=raw.device, dtype=torch.float32)
    G = torch.empty_like(raw) if want_grad else None
    L.check(L.lib().mmvqa_supcon_rows(_p(_cont(raw, "raw")), _p(mask), _p(loss_rows), _p(G), R, N, bsz, row_offset, temperature,
                                     base_temperature, _stream()), "supcon_rows")
    return loss_rows, G


def adam_step(table: Tensor, n_chunks: int, lr: float, beta1: float, beta2: float, eps: float, weight_decay: float,
              step: int, step_dev: Optional[Tensor], grad_scale: float = 1.0, max_ctas: int = 0,
              background: bool = False) -> None:
    L.check(L.lib().mmvqa_adam_step(C.cast(table.data_ptr(), C.POINTER(L.AdamDesc)), n_chunks, lr, beta1, beta2, eps,
                                   weight_decay, step, _p(step_dev), grad_scale, max_ctas, 1 if background else 0, _stream()),
            "adam_step")


def mark_rows(row_live: Tensor, ids: Tensor) -> None:
    """row_live[ids] = 1 (uint8 flags of the rows of an embedding table that receive gradient in this step)."""
    ids = ids.reshape(-1)
    if ids.dtype != torch.int64 or not ids.is_contiguous():
        ids = ids.contiguous().long()
    L.check(L.lib().mmvqa_mark_rows(_p(row_live), _p(ids), ids.numel(), row_live.numel(), _stream()), "mark_rows")


def adam_step_dev(table: Tensor, n_chunks: int, hyper_dev: Tensor, beta1: float, beta2: float, eps: float,
                  weight_decay: float, step_dev: Tensor, max_ctas: int = 0, background: bool = False) -> None:
    """Adam with lr = hyper_dev[0], grad_scale = hyper_dev[1] read on the device (graph replay follows lr schedulers)."""
    L.check(L.lib().mmvqa_adam_step_dev(C.cast(table.data_ptr(), C.POINTER(L.AdamDesc)), n_chunks, _p(hyper_dev), beta1, beta2,
                                       eps, weight_decay, _p(step_dev), max_ctas, 1 if background else 0, _stream()),
            "adam_step_dev")


def multimem_allreduce(multicast_ptr: int, signal_pads_dev: int, rank: int, world: int, nbytes: int, dtype: torch.dtype,
                       max_ctas: int = 16) -> None:
    """in-switch SUM all-reduce of a symmetric-memory bucket (csrc/comm.cu); a collective: same call on every rank."""
    L.check(L.lib().mmvqa_multimem_allreduce(multicast_ptr, signal_pads_dev, rank, world, nbytes,
                                            L.BF16 if dtype == torch.bfloat16 else L.F32, max_ctas, _stream()),
            "multimem_allreduce")


# ------------------------------------------------------------------------------- caption similarity
def jaccard_mask(ids_a: Tensor, len_a: Tensor, ids_b: Tensor, len_b: Tensor) -> Tensor:
    """ids_* [n, lmax] int32 sorted unique word ids per document, len_* [n] int32 -> [na, nb] fp32 Jaccard mask with
    ones on the diagonal (supcon_utils.py:110-138)."""
    for t in (ids_a, len_a, ids_b, len_b):
        _cont(t, "jaccard input")
        if t.dtype != torch.int32:
            raise L.MMVQAError("jaccard_mask takes int32 tensors")
    na, lmax = ids_a.shape
    nb = ids_b.shape[0]
    if ids_b.shape[1] != lmax:
        raise L.MMVQAError("jaccard_mask: both id matrices need the same row length")
    mask = torch.empty(na, nb, device=ids_a.device, dtype=torch.float32)
    L.check(L.lib().mmvqa_jaccard_mask(_p(ids_a), _p(len_a), _p(ids_b), _p(len_b), _p(mask), na, nb, lmax, _stream()),
            "jaccard_mask")
    return mask
