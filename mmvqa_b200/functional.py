"""torch.autograd.Function wrappers: each forward/backward is a short sequence of C-ABI launches.

These are the only place where autograd meets the CUDA library.  The nn.Module mirrors in
``mmvqa_b200/models`` hold the parameters (reference names/shapes) and call these functions.
All parameter gradients are produced in fp32; activations travel in the compute dtype.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import ops
from ._lib import (ACT_GELU, ACT_NONE, ACT_RELU, ACT_SERF, EPI_ACT, EPI_ACT_ROWSUM, EPI_DACT, EPI_DACT_SCALE,
                   EPI_RESIDUAL, EPI_STORE, MMVQAError)
from .config import compute_dtype

Tensor = torch.Tensor
_SM_COUNT = {}


def _sms(device) -> int:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# --------------------------------------------------------------------------------------
# side stream for work that is off the critical path (weight-gradient GEMMs): it forks from / joins the current
# stream with events, so inside a CUDA-graph capture it becomes a parallel branch of the graph
# --------------------------------------------------------------------------------------
import os as _os

_SIDE = {}
_OVERLAP = _os.environ.get("MMVQA_NO_OVERLAP") is None
_L2_PREFETCH = _os.environ.get("MMVQA_L2_PREFETCH", "1") != "0" and _OVERLAP
_LN_DEFER = _os.environ.get("MMVQA_LN_DEFER", "1") != "0"
_VISTOK_PG = _os.environ.get("MMVQA_VISTOK_PG", "1") != "0"
TRAIN_STEP_CAPTURE = False      # set by graph.GraphedTrainStep while it captures forward + backward
_EMBED_PREZERO = _os.environ.get("MMVQA_EMBED_PREZERO", "1") != "0"
_WGRAD_LATE = _os.environ.get("MMVQA_WGRAD_LATE", "0") == "1"   # measured neutral (2.455 vs 2.457 ms/step; hot path 2.194 vs 2.158): opt-in
_L2_PREFETCH_CTAS = int(_os.environ.get("MMVQA_L2_PREFETCH_CTAS", "16"))


class SideBranch:
    """`with branch.after_now(): ...` runs the body on the side stream once everything enqueued so far on the main
    stream has finished; `join()` makes the main stream wait for the branch.  Tensors touched by the branch must be
    kept alive by the caller until join() (the caching allocator only tracks the allocating stream).
    `index` selects one of several side streams (independent branches that may all run at once)."""

    def __init__(self, device, index: int = 0, priority: int = 0):
        self.main = torch.cuda.current_stream(device)
        key = (device.index if device.index is not None else torch.cuda.current_device(), self.main.cuda_stream, index,
               priority)
        if key not in _SIDE:
            # priority -1: same class as a high-priority main stream (graph.GraphedTrainStep main_priority); the default
            # 0 yields the SMs to the main chain (weight-gradient GEMMs, optimizer)
            _SIDE[key] = torch.cuda.Stream(device, priority=priority)
        self.side = _SIDE[key] if _OVERLAP else self.main
        self.used = False

    def after_now(self):
        if self.side is not self.main:
            ev = torch.cuda.Event()
            ev.record(self.main)
            self.side.wait_event(ev)
            self.used = True
        return torch.cuda.stream(self.side)

    def join(self):
        if self.used:
            ev = torch.cuda.Event()
            ev.record(self.side)
            self.main.wait_event(ev)
            self.used = False


# --------------------------------------------------------------------------------------
# gradient sink: lets an optimizer / data-parallel engine consume parameter gradients the moment a layer's backward
# has enqueued them, instead of after the whole backward (autograd only publishes .grad when a node returns)
# --------------------------------------------------------------------------------------
_GRAD_SINK = None


def set_grad_sink(fn) -> None:
    """``fn(params, grads, side_stream, tail) -> bool`` is called from inside multi-layer backward nodes with the
    parameters of ONE layer and their final fp32 gradients; work producing them is enqueued on the current stream
    and on `side_stream`; `tail` is True for the last layer of the node (no backward work is left to overlap with).  Returning True means the sink took the gradients (it may set ``p.grad`` itself): the node
    then returns None for them.  ``None`` removes the sink."""
    global _GRAD_SINK
    _GRAD_SINK = fn


def grad_sink():
    return _GRAD_SINK


# --------------------------------------------------------------------------------------
# weight cache: fp32 master parameters -> compute-dtype GEMM operands (optionally concatenated)
# --------------------------------------------------------------------------------------
class WeightCache:
    """Casts parameters to the compute dtype once per parameter version.

    The key is the parameter identity; an entry is valid while every source parameter keeps its ``_version`` and
    storage and no optimizer that bypasses torch's version counters has stepped since (``epoch``).
    ``torch.optim`` steps and ``load_state_dict`` bump the version; the fused Adam (mmvqa_b200.optim) writes the
    bf16 copies itself in its update kernel and re-validates exactly the entries it refreshed."""

    def __init__(self):
        self._entries = {}
        self.epoch = 0
        self.generation = 0          # bumps whenever a buffer is (re)allocated or the cache is cleared

    def _sig(self, params):
        return (self.epoch, tuple((p._version, p.data_ptr()) for p in params))

    def get(self, params: Sequence[Tensor], dtype: torch.dtype) -> Tensor:
        key = (tuple(id(p) for p in params), dtype)
        sig = self._sig(params)
        ent = self._entries.get(key)
        if ent is not None and ent[0] == sig:
            return ent[1]
        with torch.no_grad():
            flat = [p.detach().reshape(p.shape[0], -1) for p in params]
            if dtype == torch.float32 and len(flat) == 1:
                out = flat[0]
            else:
                rows = sum(f.shape[0] for f in flat)
                if ent is not None and ent[1].shape == (rows, flat[0].shape[1]):
                    out = ent[1]
                else:
                    out = torch.empty(rows, flat[0].shape[1], device=flat[0].device, dtype=dtype)
                    self.generation += 1
                r = 0
                for f in flat:
                    ops.cast(f.contiguous(), dtype, out=out[r:r + f.shape[0]])
                    r += f.shape[0]
        self._entries[key] = (sig, out, list(params))
        return out

    def bf16_view(self, p: Tensor) -> Optional[Tensor]:
        """The bf16 copy of `p` inside the cache (a row slice of a concatenated entry if need be), or None.
        Returns None as well if `p` appears in more than one bf16 entry (then the version check recasts)."""
        found = None
        for (ids, dtype), (_, out, params) in self._entries.items():
            if dtype != torch.bfloat16 or id(p) not in ids:
                continue
            r = 0
            for q in params:
                if q is p:
                    if found is not None:
                        return None
                    found = out[r:r + q.shape[0]]
                r += q.shape[0]
        return found

    def note_optimizer_step(self, refreshed_ids) -> None:
        """An optimizer updated parameters behind torch's back: entries whose parameters were all refreshed in
        the same kernel stay valid, every other entry is recast on next use."""
        self.epoch += 1
        for key, (sig, out, params) in list(self._entries.items()):
            if key[1] == torch.bfloat16 and all(id(q) in refreshed_ids for q in params):
                self._entries[key] = (self._sig(params), out, params)

    def covered_by(self, refreshed_ids) -> bool:
        return all(all(id(q) in refreshed_ids for q in params) for (ids, dtype), (_, _, params) in self._entries.items()
                   if not (dtype == torch.float32 and len(params) == 1))

    def clear(self) -> None:
        self._entries.clear()
        self.generation += 1


weight_cache = WeightCache()


def invalidate_weight_cache() -> None:
    weight_cache.clear()


# --------------------------------------------------------------------------------------
# GEMM helpers (all row-major 2-D views)
# --------------------------------------------------------------------------------------
def _pad_ld(t2d: Tensor, dtype: torch.dtype) -> Tuple[Tensor, int]:
    """Return (tensor, ld) with the compute dtype and, for bf16, a leading dimension that is a
    multiple of 8 elements (TMA needs 16-byte row strides)."""
    rows, cols = t2d.shape
    need = 8 if dtype == torch.bfloat16 else 1
    if t2d.dtype == dtype and t2d.is_contiguous() and cols % need == 0 and t2d.data_ptr() % 16 == 0:
        return t2d, cols
    t2d = t2d.contiguous()
    ld = _round_up(cols, need)
    return ops.cast_pad(t2d, rows, cols, cols, dtype, ld), ld


def _split_k_for(tiles: int, k: int, device, kblock: int = 64) -> int:
    sms = _sms(device)
    if tiles >= sms:
        return 1
    kblocks = max(1, (k + kblock - 1) // kblock)
    return max(1, min(sms // max(tiles, 1), kblocks // 4 if kblocks >= 8 else 1))


def split_k_slabs(M: int, N: int, K: int, device, dtype: torch.dtype) -> int:
    """How many fp32 split-K slabs a [M, N] = [M, K] x [K, N] GEMM should write so that it covers the chip (1 = the
    plain fused-epilogue GEMM).  Small-batch problems only: with M = 448 a K = 3072 contraction has 48 output tiles
    for 148 SMs, and each CTA would pull its whole K range through one SM's L2 port."""
    if dtype != torch.bfloat16 or N % 8 != 0 or N > 1024:
        return 1
    sms = _sms(device)
    tiles = ((M + 127) // 128) * ((N + 63) // 64)
    kblocks = (K + 63) // 64
    if tiles * 2 > sms or kblocks < 24:
        return 1
    ns = min(sms // tiles, kblocks // 8)
    while ns > 1 and kblocks % ns:
        ns -= 1
    return max(ns, 1)


def gemm_dgrad(dy: Tensor, ld_dy: int, M: int, N: int, w: Tensor, K: int, *, epilogue=EPI_STORE, act=ACT_NONE,
               aux_in: Optional[Tensor] = None, out_dtype: Optional[torch.dtype] = None,
               colsum_out: Optional[Tensor] = None) -> Tensor:
    """dx[M,K] = dy[M,N] . w[N,K]   (w read MN-major: no transposed weight copy).  colsum_out (zero-filled fp32 [K])
    receives the column sums of dx: the bias gradient of the layer that produced this activation."""
    dx = torch.empty(M, K, device=dy.device, dtype=out_dtype or dy.dtype)
    ops.gemm(M, K, N, dy, ld_dy, False, w, K, True, dx, K, epilogue=epilogue, act=act, aux_in=aux_in, ld_aux_in=K,
             colsum_out=colsum_out, b_static=True)
    return dx


def gemm_wgrad(dy: Tensor, ld_dy: int, M: int, N: int, x: Tensor, ld_x: int, K: int, zeroed: Optional[Tensor] = None) -> Tensor:
    """dW[N,K] (fp32) = dy[M,N]^T . x[M,K]   (both operands read MN-major); split-K when the output has
    too few tiles to fill the chip.  `zeroed`: an already zero-filled [N,K] fp32 buffer to accumulate into."""
    tile = 128 if dy.dtype == torch.bfloat16 else 64
    tiles = ((N + tile - 1) // tile) * ((K + tile - 1) // tile)
    sk = _split_k_for(tiles, M, dy.device, 64 if dy.dtype == torch.bfloat16 else 16)
    if sk > 1:
        dw = zeroed if zeroed is not None else torch.zeros(N, K, device=dy.device, dtype=torch.float32)
        ops.gemm(N, K, M, dy, ld_dy, True, x, ld_x, True, dw, K, accumulate=True, split_k=sk, b_static=True)
    else:
        dw = torch.empty(N, K, device=dy.device, dtype=torch.float32)
        ops.gemm(N, K, M, dy, ld_dy, True, x, ld_x, True, dw, K, b_static=True)
    return dw


# --------------------------------------------------------------------------------------
# casts
# --------------------------------------------------------------------------------------
class CastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, dtype: torch.dtype):
        ctx.src_dtype = x.dtype
        return ops.cast(x.contiguous(), dtype)

    @staticmethod
    def backward(ctx, dy):
        return ops.cast(dy.contiguous(), ctx.src_dtype), None


def to_compute(x: Tensor) -> Tensor:
    dt = compute_dtype()
    if x.dtype == dt:
        return x if x.is_contiguous() else x.contiguous()
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return CastFn.apply(x, dt)


def to_float32(x: Tensor) -> Tensor:
    if x.dtype == torch.float32:
        return x
    return CastFn.apply(x, torch.float32)


# --------------------------------------------------------------------------------------
# Linear (+ activation | + residual with dropout)
# --------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = epilogue(x W^T + b).  act != NONE -> y = act(.) ; residual != None -> y = dropout(.) + residual."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], act: int, out_fp32: bool,
                residual: Optional[Tensor], dropout_p: float, seed: int):
        dt = x.dtype
        w = weight_cache.get((weight,), dt)
        N, K = w.shape
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M = x2.shape[0]
        out_dt = torch.float32 if out_fp32 else dt
        y = torch.empty(M, N, device=x.device, dtype=out_dt)
        pre = None
        b = None if bias is None else bias.detach()
        if act != ACT_NONE:
            if residual is not None:
                raise MMVQAError("LinearFn: activation and residual are exclusive")
            pre = torch.empty(M, N, device=x.device, dtype=dt)
            ops.gemm(M, N, K, x2, K, False, w, K, False, y, N, bias=b, epilogue=EPI_ACT, act=act, aux_out=pre, ld_aux_out=N,
                     b_static=True)
        elif residual is not None:
            r2 = residual.reshape(M, N)
            if not r2.is_contiguous():
                r2 = r2.contiguous()
            ops.gemm(M, N, K, x2, K, False, w, K, False, y, N, bias=b, epilogue=EPI_RESIDUAL, aux_in=r2, ld_aux_in=N,
                     dropout_p=dropout_p, dropout_seed=seed, b_static=True)
        else:
            ops.gemm(M, N, K, x2, K, False, w, K, False, y, N, bias=b, b_static=True)
        ctx.save_for_backward(x2, weight, pre)
        ctx.meta = (act, dropout_p, seed, bias is not None, residual is not None, x.shape, weight.shape, dt)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy: Tensor):
        x2, weight, pre = ctx.saved_tensors
        act, p, seed, has_bias, has_res, xshape, wshape, dt = ctx.meta
        w = weight_cache.get((weight,), dt)
        N, K = w.shape
        M = x2.shape[0]
        dy2 = dy.reshape(M, N)
        dres = None
        if has_res and ctx.needs_input_grad[5]:
            dres = dy2.to(dt).view(*xshape[:-1], N) if dy2.dtype != dt else dy2.view(*xshape[:-1], N)
        g, ld = _pad_ld(dy2, dt)
        if has_res and p > 0.0:
            if ld != N:
                raise MMVQAError("dropout backward needs an 8-aligned width")
            g = ops.dropout(g, p, seed)
        if act != ACT_NONE:
            if ld != N:
                g = g[:, :N].contiguous()
                ld = N
            g = ops.bias_act_bwd(pre, None, g, act)
            g, ld = _pad_ld(g, dt)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm_dgrad(g, ld, M, N, w, K).view(xshape)
        if ctx.needs_input_grad[1]:
            dw = gemm_wgrad(g, ld, M, N, x2, K, K).view(wshape)
        if has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum(g, M, N, ld)
        return dx, dw, db, None, None, dres, None, None


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, *, act: int = ACT_NONE, out_fp32: bool = False,
           residual: Optional[Tensor] = None, dropout_p: float = 0.0, seed: int = 0) -> Tensor:
    return LinearFn.apply(x, weight, bias, act, out_fp32, residual, dropout_p, seed)


# --------------------------------------------------------------------------------------
# standalone bias + activation (SERF module, gelu())
# --------------------------------------------------------------------------------------
class ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, act: int):
        x = x.contiguous()
        ctx.save_for_backward(x)
        ctx.act = act
        return ops.bias_act_fwd(x, None, act)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.bias_act_bwd(x, None, dy.contiguous().to(x.dtype), ctx.act), None


# --------------------------------------------------------------------------------------
# residual + LayerNorm
# --------------------------------------------------------------------------------------
class AddLayerNormFn(torch.autograd.Function):
    """y = LN(x + res) * gamma + beta   (res may be None)."""

    @staticmethod
    def forward(ctx, x: Tensor, res: Optional[Tensor], gamma: Tensor, beta: Tensor, eps: float):
        x = x.contiguous()
        r = None if res is None else res.contiguous()
        y, xsum, mean, rstd = ops.add_layernorm_fwd(x, r, gamma.detach(), beta.detach(), eps, want_sum=r is not None)
        ctx.save_for_backward(x if xsum is None else xsum, gamma, mean, rstd)
        ctx.has_res = res is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        xsum, gamma, mean, rstd = ctx.saved_tensors
        dg = torch.zeros_like(gamma, dtype=torch.float32)
        db = torch.zeros_like(gamma, dtype=torch.float32)
        dx = ops.layernorm_bwd(dy.contiguous().to(xsum.dtype), xsum, gamma.detach(), mean, rstd, None, dg, db)
        return dx, (dx if ctx.has_res else None), dg, db, None


def add_layer_norm(x: Tensor, res: Optional[Tensor], gamma: Tensor, beta: Tensor, eps: float) -> Tensor:
    return AddLayerNormFn.apply(x, res, gamma, beta, eps)


# --------------------------------------------------------------------------------------
# pooling / normalisation
# --------------------------------------------------------------------------------------
class MaskedMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h: Tensor, maskf: Tensor):
        ctx.save_for_backward(maskf)
        ctx.T = h.shape[1]
        return ops.masked_mean_fwd(h.contiguous(), maskf)

    @staticmethod
    def backward(ctx, dout):
        (maskf,) = ctx.saved_tensors
        return ops.masked_mean_bwd(dout.contiguous(), maskf, ctx.T), None


class L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor):
        y, inv = ops.l2norm_fwd(x.contiguous())
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        return ops.l2norm_bwd(y, inv, dy.contiguous())


# --------------------------------------------------------------------------------------
# visual-token projector: v[b,:] = mean_hw act(W . f[b,:,hw])     (image_encoding.py:74,103)
# --------------------------------------------------------------------------------------
class VisTokFn(torch.autograd.Function):
    """One pyramid level.  The [B, hidden, H, W] activation map is never materialised; when a gradient is needed the
    forward epilogue also stores act'(.) (compute dtype), so backward is two plain GEMMs with no transcendental:
    dW = sum_b (dv_b / HW) . (act'_b f_b^T)   and   df_b = (W * dv_b / HW)^T act'_b."""

    @staticmethod
    def forward(ctx, feat: Tensor, conv_w: Tensor, act: int, dtype: torch.dtype):
        B, Cc, Hh, Ww = feat.shape
        HW = Hh * Ww
        w = weight_cache.get((conv_w,), dtype)            # [hidden, C]
        hidden = w.shape[0]
        f2 = feat.detach().reshape(B * Cc, HW)
        fb, ld = _pad_ld(f2, dtype)                         # compute dtype, ld % 8 == 0 for bf16
        v = torch.zeros(B, hidden, device=feat.device, dtype=torch.float32)
        need_bwd = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        actp = torch.empty(B, hidden, ld, device=feat.device, dtype=dtype) if need_bwd else None
        ops.gemm(hidden, HW, Cc, w, Cc, False, fb, ld, True, None, 0, epilogue=EPI_ACT_ROWSUM, act=act, rowsum_out=v,
                 scale=1.0 / HW, batch=B, a_batch_rows=0, b_batch_rows=Cc, aux_out=actp, ld_aux_out=ld)
        ctx.save_for_backward(fb, conv_w, actp)
        ctx.meta = (act, dtype, B, Cc, Hh, Ww, ld, hidden, feat.dtype)
        return v

    @staticmethod
    def backward(ctx, dv: Tensor):
        fb, conv_w, actp = ctx.saved_tensors
        act, dtype, B, Cc, Hh, Ww, ld, hidden, fdt = ctx.meta
        HW = Hh * Ww
        dvs = dv.contiguous().float()
        dw = dfeat = None
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(hidden, Cc, device=fb.device, dtype=torch.float32)
            tiles = ((hidden + 127) // 128) * B
            kblocks = (HW + 63) // 64
            sk = max(1, min(4, (2 * _sms(fb.device)) // max(tiles, 1), kblocks // 16))
            ops.gemm(hidden, Cc, HW, actp, ld, False, fb, ld, False, dw, Cc, accumulate=True, split_k=sk, batch=B,
                     a_batch_rows=hidden, b_batch_rows=Cc, c_batch_stride=0, rowscale=dvs, scale=1.0 / HW)
            dw = dw.view(conv_w.shape)
        if ctx.needs_input_grad[0]:
            w32 = conv_w.detach().reshape(hidden, Cc).float()
            wb = (w32.unsqueeze(0) * (dvs / HW).unsqueeze(-1)).to(dtype).contiguous()       # [B, hidden, C]
            dfeat = torch.empty(B, Cc, HW, device=fb.device, dtype=torch.float32)
            ops.gemm(Cc, HW, hidden, wb, Cc, True, actp, ld, True, dfeat, HW, batch=B, a_batch_rows=hidden,
                     b_batch_rows=hidden, c_batch_stride=Cc * HW)
            dfeat = dfeat.view(B, Cc, Hh, Ww).to(fdt)
        return dfeat, dw, None, None


def vistok_project(feat: Tensor, conv_w: Tensor, act: int) -> Tensor:
    if feat.dtype not in (torch.float32, torch.bfloat16):
        feat = feat.float()
    return VisTokFn.apply(feat, conv_w, act, compute_dtype())


class VisTokAllFn(torch.autograd.Function):
    """All pyramid levels in one autograd node: v[n] = mean_hw act(W_n . f_n).  The levels are independent, so the
    largest one runs on the main stream and the others on a side branch (parallel graph branches when captured);
    returns the stacked [num_vis, B, hidden] fp32 tokens that the fused embedding kernel consumes."""

    @staticmethod
    def forward(ctx, act: int, dtype: torch.dtype, nlev: int, *tensors):
        feats, convs = tensors[:nlev], tensors[nlev:]
        B = feats[0].shape[0]
        hidden = convs[0].shape[0]
        dev = feats[0].device
        vis = torch.zeros(nlev, B, hidden, device=dev, dtype=torch.float32)
        order = sorted(range(nlev), key=lambda n: -feats[n][0].numel())          # biggest level first, on main
        branch = SideBranch(dev, priority=-1)            # on the critical path like the main chain
        saved, metas = [None] * (3 * nlev), [None] * nlev
        keep = []

        # operand casts of every level first, on the main stream: a cast issued next to another level's projector kernel
        # waits for SM slots (67 us instead of 13 for the 112 x 112 level) and it heads the critical path
        # (fp32 maps -> bf16: ONE launch for all levels instead of five)
        flat = [f.detach().reshape(B * f.shape[1], f.shape[2] * f.shape[3]) for f in feats]
        if dtype == torch.bfloat16 and all(t.dtype in (torch.float32, torch.bfloat16) for t in flat):
            # levels that are already bf16 with 16-byte rows are the TMA operand as they are; the rest share one launch
            inplace = [t.dtype == torch.bfloat16 and t.is_contiguous() and t.shape[1] % 8 == 0 and t.data_ptr() % 16 == 0
                       for t in flat]
            todo = [n for n in range(nlev) if not inplace[n]]
            lds = [_round_up(t.shape[1], 8) for t in flat]
            done = ops.cast_pad_multi([flat[n].contiguous() for n in todo], [lds[n] for n in todo]) if todo else []
            casted = [(flat[n], lds[n]) for n in range(nlev)]
            for n, t in zip(todo, done):
                casted[n] = (t, lds[n])
        else:
            casted = [_pad_ld(t, dtype) for t in flat]

        def run(n):
            f, cw = feats[n], convs[n]
            _, Cc, Hh, Ww = f.shape
            HW = Hh * Ww
            w = weight_cache.get((cw,), dtype)
            fb, ld = casted[n]
            need_df, need_dw = ctx.needs_input_grad[3 + n], ctx.needs_input_grad[3 + nlev + n]
            if _VISTOK_PG and dtype == torch.bfloat16 and need_dw and not need_df and ops.vistok_pgrad_supported(hidden, HW, Cc):
                # only the weight gradient is needed: the pixel contraction of act' with the map is finished inside the
                # forward kernel (P [B, hidden, C]); the [B, hidden, HW] act' map is never written
                pg = ops.vistok_fwd_pgrad(w.reshape(hidden, Cc), fb, ld, vis[n], B, hidden, HW, Cc, act)
                saved[3 * n:3 * n + 3] = [None, cw, pg]
                metas[n] = (Cc, Hh, Ww, ld, f.dtype, True)
                keep.append((fb, pg))
                return
            actp = torch.empty(B, hidden, ld, device=dev, dtype=dtype) if (need_df or need_dw) else None
            ops.gemm(hidden, HW, Cc, w, Cc, False, fb, ld, True, None, 0, epilogue=EPI_ACT_ROWSUM, act=act, rowsum_out=vis[n],
                     scale=1.0 / HW, batch=B, a_batch_rows=0, b_batch_rows=Cc, aux_out=actp, ld_aux_out=ld)
            saved[3 * n:3 * n + 3] = [fb, cw, actp]
            metas[n] = (Cc, Hh, Ww, ld, f.dtype, False)
            keep.append((fb, actp))
        # fork BEFORE the big level is enqueued: the side branch only depends on what precedes this node.  (Issuing the
        # big level first was measured slower: the small levels then queue behind it and finish 45 us later.)
        # (one branch per level, as in the backward pass: the small levels are latency-bound launches that would otherwise
        # queue behind each other on one stream and end after the big level)
        branches = [branch] + [SideBranch(dev, index=1 + i, priority=-1) for i in range(1, max(nlev - 1, 0))]
        for br, n in zip(branches, order[1:]):
            with br.after_now():
                run(n)
        run(order[0])
        for br in branches:
            br.join()
        ctx.save_for_backward(*saved)
        ctx.meta = (act, dtype, nlev, B, hidden, metas, order)
        return vis

    @staticmethod
    def backward(ctx, dvis: Tensor):
        act, dtype, nlev, B, hidden, metas, order = ctx.meta
        saved = ctx.saved_tensors
        dev = dvis.device
        dvs = dvis.contiguous().float()
        dfeats, dws = [None] * nlev, [None] * nlev
        keep = []

        def run(n):
            fb, cw, actp = saved[3 * n:3 * n + 3]
            Cc, Hh, Ww, ld, fdt, is_pg = metas[n]
            HW = Hh * Ww
            dv = dvs[n]
            if is_pg:           # actp holds P = sum_hw act' f: the weight gradient is a [hidden, C] contraction over the batch
                dws[n] = ops.vistok_dw(actp, dv, 1.0 / HW).view(cw.shape)
                return
            if ctx.needs_input_grad[3 + nlev + n]:
                dw = torch.zeros(hidden, Cc, device=dev, dtype=torch.float32)
                tiles = ((hidden + 127) // 128) * B
                kblocks = (HW + 63) // 64
                sk = max(1, min(4, (2 * _sms(dev)) // max(tiles, 1), kblocks // 16))
                ops.gemm(hidden, Cc, HW, actp, ld, False, fb, ld, False, dw, Cc, accumulate=True, split_k=sk, batch=B,
                         a_batch_rows=hidden, b_batch_rows=Cc, c_batch_stride=0, rowscale=dv, scale=1.0 / HW)
                dws[n] = dw.view(cw.shape)
            if ctx.needs_input_grad[3 + n]:
                w32 = cw.detach().reshape(hidden, Cc).float()
                wb = (w32.unsqueeze(0) * (dv / HW).unsqueeze(-1)).to(dtype).contiguous()
                df = torch.empty(B, Cc, HW, device=dev, dtype=torch.float32)
                ops.gemm(Cc, HW, hidden, wb, Cc, True, actp, ld, True, df, HW, batch=B, a_batch_rows=hidden, b_batch_rows=hidden,
                         c_batch_stride=Cc * HW)
                dfeats[n] = df.view(B, Cc, Hh, Ww).to(fdt)
                keep.append(wb)
        # every level is an independent pair of small GEMMs (atomics / latency bound): one branch per level
        branches = [SideBranch(dev, index=1 + i, priority=-1) for i in range(max(nlev - 1, 0))]
        for br, n in zip(branches, order[1:]):
            with br.after_now():
                run(n)
        run(order[0])
        for br in branches:
            br.join()
        return (None, None, None, *dfeats, *dws)


def vistok_project_all(feats: Sequence[Tensor], conv_ws: Sequence[Tensor], act: int) -> Tensor:
    """[num_vis, B, hidden] fp32 visual tokens."""
    fs = [f if f.dtype in (torch.float32, torch.bfloat16) else f.float() for f in feats]
    return VisTokAllFn.apply(act, compute_dtype(), len(fs), *fs, *conv_ws)


# --------------------------------------------------------------------------------------
# BertEmbeddings + visual-token scatter  (mmbert.py:60-67)
# --------------------------------------------------------------------------------------
class EmbedFuseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, seg, word, pos, typ, gamma, beta, vis, eps, p, seed, out_dtype, padding_idx):
        ids = ids.contiguous().long()
        seg = seg.contiguous().long()
        visc = None if vis is None else vis.contiguous().float()
        h, mean, rstd = ops.embed_ln_scatter_fwd(ids, seg, word.detach(), pos.detach(), typ.detach(), gamma.detach(),
                                                 beta.detach(), visc, out_dtype, eps, p, seed)
        ctx.save_for_backward(ids, seg, word, pos, typ, gamma, mean, rstd)
        ctx.meta = (0 if vis is None else vis.shape[0], p, seed, padding_idx)
        # the dense word-table gradient (94 MB for bert-base) must be zero-filled before the backward kernel scatters into
        # it: 20 us at the very end of the step's critical path.  Fill it now, on a side branch under the encoder.
        ctx.prezero = None
        # (inside a stream capture only when the capture is known to contain the backward pass -- graph.GraphedTrainStep
        # sets TRAIN_STEP_CAPTURE -- because the branch is joined there: a forward-only capture must not end with it open)
        if _EMBED_PREZERO and ctx.needs_input_grad[2] and word.numel() >= (1 << 20) and _OVERLAP and \
                (TRAIN_STEP_CAPTURE or not torch.cuda.is_current_stream_capturing()):
            br = SideBranch(word.device, index=8)
            with br.after_now():
                ctx.prezero = (torch.zeros_like(word, dtype=torch.float32), br)
        return h

    @staticmethod
    def backward(ctx, dh):
        ids, seg, word, pos, typ, gamma, mean, rstd = ctx.saved_tensors
        nvis, p, seed, padding_idx = ctx.meta
        B, T = ids.shape
        H = word.shape[1]
        need = ctx.needs_input_grad
        dword = None
        if need[2]:
            if ctx.prezero is not None:
                dword, br = ctx.prezero
                ctx.prezero = None          # a second backward through this node gets a fresh buffer
                br.join()
            else:
                dword = torch.zeros_like(word, dtype=torch.float32)
        dpos = torch.zeros_like(pos, dtype=torch.float32) if need[3] else None
        dtyp = torch.zeros_like(typ, dtype=torch.float32) if need[4] else None
        dgamma = torch.zeros_like(gamma, dtype=torch.float32) if need[5] else None
        dbeta = torch.zeros_like(gamma, dtype=torch.float32) if need[6] else None
        dvis = torch.empty(nvis, B, H, device=dh.device, dtype=torch.float32) if (nvis > 0 and need[7]) else None
        ops.embed_ln_scatter_bwd(dh.contiguous(), ids, seg, word.detach(), pos.detach(), typ.detach(), gamma.detach(), mean,
                                 rstd, dword, dpos, dtyp, dgamma, dbeta, dvis, nvis, padding_idx, p, seed)
        return None, None, dword, dpos, dtyp, dgamma, dbeta, dvis, None, None, None, None, None


# --------------------------------------------------------------------------------------
# Transformer multi-head self-attention core (fused QKV projection + attention)
# --------------------------------------------------------------------------------------
class MHSAFn(torch.autograd.Function):
    """h = merge_heads(softmax(q k^T / sqrt(d) - 10000 (1 - mask_j)) v) with q,k,v = x W^T + b from ONE
    [M,H] x [H,3H] GEMM.  Returns (h, probs)."""

    @staticmethod
    def forward(ctx, x, maskf, wq, bq, wk, bk, wv, bv, heads, p, seed):
        dt = x.dtype
        B, T, H = x.shape
        d = H // heads
        w = weight_cache.get((wq, wk, wv), dt)                       # [3H, H]
        bias = torch.cat([bq.detach(), bk.detach(), bv.detach()]).float()
        x2 = x.reshape(B * T, H)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        qkv = torch.empty(B * T, 3 * H, device=x.device, dtype=dt)
        ops.gemm(B * T, 3 * H, H, x2, H, False, w, H, False, qkv, 3 * H, bias=bias, b_static=True)
        out, probs = ops.mhsa_fwd(qkv, maskf, B, T, heads, d, p, seed)
        ctx.save_for_backward(x2, qkv, probs, wq, wk, wv)
        ctx.meta = (B, T, H, heads, d, p, seed, dt)
        ctx.mark_non_differentiable(probs)
        return out.view(B, T, H), probs

    @staticmethod
    def backward(ctx, dout, _dprobs):
        x2, qkv, probs, wq, wk, wv = ctx.saved_tensors
        B, T, H, heads, d, p, seed, dt = ctx.meta
        w = weight_cache.get((wq, wk, wv), dt)
        M = B * T
        do2 = dout.reshape(M, H).contiguous().to(dt)
        dqkv = ops.mhsa_bwd(qkv, probs, do2, B, T, heads, d, p, seed)
        dx = gemm_dgrad(dqkv, 3 * H, M, 3 * H, w, H).view(B, T, H) if ctx.needs_input_grad[0] else None
        dw = gemm_wgrad(dqkv, 3 * H, M, 3 * H, x2, H, H)
        db = ops.colsum(dqkv, M, 3 * H)
        return (dx, None, dw[:H], db[:H], dw[H:2 * H], db[H:2 * H], dw[2 * H:], db[2 * H:], None, None, None)


class DropoutFn(torch.autograd.Function):
    """nn.Dropout with the library's counter-hash mask (same mask regenerated in backward)."""

    @staticmethod
    def forward(ctx, x: Tensor, p: float, seed: int):
        ctx.meta = (p, seed)
        return ops.dropout(x.contiguous(), p, seed)

    @staticmethod
    def backward(ctx, dy):
        p, seed = ctx.meta
        return ops.dropout(dy.contiguous(), p, seed), None, None


class RFAttentionFn(torch.autograd.Function):
    """RealFormer attention sub-block without the out-projection (realformer.py:30-44):
    shared-weight kqv GEMM + fused residual attention.  Returns (merged heads [B,T,H], scores [B,h,T,T])."""

    @staticmethod
    def forward(ctx, x, maskf, prev, kqv_w, heads):
        ctx.set_materialize_grads(False)
        dt = x.dtype
        B, T, H = x.shape
        d = H // heads
        M = B * T
        wk = weight_cache.get((kqv_w,), dt)
        xin = x.reshape(M, H)
        if not xin.is_contiguous():
            xin = xin.contiguous()
        kqv = torch.empty(M * heads, 3 * d, device=x.device, dtype=dt)
        ops.gemm(M * heads, 3 * d, d, xin, d, False, wk, d, False, kqv, 3 * d, b_static=True)
        prevc = None if prev is None else prev.contiguous().float()
        attn, scores = ops.rf_attn_fwd(kqv, prevc, maskf, B, T, heads, d)
        ctx.save_for_backward(xin, kqv, scores, kqv_w)
        ctx.meta = (B, T, H, heads, d, dt, prev is not None)
        return attn.view(B, T, H), scores

    @staticmethod
    def backward(ctx, dattn, dscores):
        xin, kqv, scores, kqv_w = ctx.saved_tensors
        B, T, H, heads, d, dt, has_prev = ctx.meta
        M = B * T
        wk = weight_cache.get((kqv_w,), dt)
        if dattn is None:
            da = torch.zeros(M, H, device=xin.device, dtype=dt)
        else:
            da = dattn.reshape(M, H).contiguous().to(dt)
        ds = None if dscores is None else dscores.contiguous().float()
        want_dprev = has_prev and ctx.needs_input_grad[2]
        dkqv, dprev = ops.rf_attn_bwd(kqv, scores, da, ds, want_dprev, B, T, heads, d)
        dwk = gemm_wgrad(dkqv, 3 * d, M * heads, 3 * d, xin, d, d).view(kqv_w.shape)
        dx = torch.empty(M, H, device=xin.device, dtype=dt)
        ops.gemm(M * heads, d, 3 * d, dkqv, 3 * d, False, wk, d, True, dx, d, b_static=True)
        return dx.view(B, T, H), None, dprev, dwk, None


# --------------------------------------------------------------------------------------
# RealFormer encoder: L post-LN residual-attention blocks in one autograd node
# --------------------------------------------------------------------------------------
RF_PARAMS_PER_LAYER = 10  # kqv.w proj.w ln1.w ln1.b ff.0.w ff.0.b ff.2.w ff.2.b ln2.w ln2.b


class RealFormerEncoderFn(torch.autograd.Function):
    """models/realformer.py:30-51 x n_layers (mmbert.py:103-108), 7 launches per layer forward:
    kqv GEMM | fused residual attention (prev in, scores out) | proj GEMM + dropout + residual |
    LN1 | FF1 GEMM + bias + SERF | FF2 GEMM + bias + dropout + residual | LN2."""

    @staticmethod
    def forward(ctx, x, maskf, prev, heads, p1, p2, seed, *params):
        ctx.set_materialize_grads(False)
        dt = x.dtype
        B, T, H = x.shape
        d = H // heads
        M = B * T
        n_layers = len(params) // RF_PARAMS_PER_LAYER
        xin = x.reshape(M, H)
        if not xin.is_contiguous():
            xin = xin.contiguous()
        if prev is not None:
            prev = prev.contiguous().float()
        saved: List[Tensor] = []
        scores = prev
        parts = None
        # the kqv projection rides inside the attention kernel on the tensor-core path (one launch less per layer)
        fuse_kqv = (dt == torch.bfloat16 and d % 16 == 0 and T <= 128 and d <= 128 and
                    (2 * ((T + 15) // 16 * 16) * (d + 8) + d * ((T + 15) // 16 * 16 + 8) + ((T + 15) // 16 * 16) * (d + 8)
                     + 3 * d * (d + 8)) * 2 <= 200 * 1024 and _os.environ.get("MMVQA_NO_FUSED_KQV") is None)
        # opt-in (MMVQA_RF_ENCODER=1): the whole encoder forward as ONE launch of the sample-stationary cluster kernel
        # (csrc/rf_encoder.cu); it writes the same intermediates the per-operator chain below saves.  Measured at the
        # flagship shape (B = 16, T = 28, 12 layers): 717 us against 634 us for the chain below -- one 8-CTA cluster per
        # sample pair keeps 64 of the 148 SMs busy and each of them is bound by its own TMA ingest (~0.7 us per 48 KB box
        # with two boxes in flight), so the chain, which spreads every GEMM over the whole chip, stays the default.
        F4_ = params[4].shape[0]
        if dt == torch.bfloat16 and _os.environ.get("MMVQA_RF_ENCODER") == "1" and \
                ops.rf_encoder_supported(B, T, H, heads, F4_, n_layers):
            layers = []
            for l in range(n_layers):
                kqv_w, proj_w, g1, b1, w0, bb0, w2, bb2, g2, b2 = params[l * RF_PARAMS_PER_LAYER:(l + 1) * RF_PARAMS_PER_LAYER]
                layers.append((weight_cache.get((kqv_w,), dt), weight_cache.get((proj_w,), dt), weight_cache.get((w0,), dt),
                               weight_cache.get((w2,), dt), bb0.detach(), bb2.detach(), g1.detach(), b1.detach(), g2.detach(),
                               b2.detach()))
            o = ops.rf_encoder_fwd(xin, layers, prev, maskf, B, T, heads, p1, p2, 1e-5, seed)
            for l in range(n_layers):
                saved += [xin if l == 0 else o["xout"][l - 1], o["kqv"][l], o["scores"][l], o["att"][l], o["y1"][l],
                          o["mean1"][l], o["rstd1"][l], o["x1"][l], o["hpre"][l], o["hact"][l], o["y2"][l], o["mean2"][l],
                          o["rstd2"][l]]
            ctx.save_for_backward(*saved, *params)
            ctx.meta = (B, T, H, heads, d, n_layers, p1, p2, seed, dt, prev is not None)
            return o["xout"][n_layers - 1].view(B, T, H), o["scores"][n_layers - 1]
        # L2 prefetch one layer ahead (side branch, hints only): 12 layers x 14 MB of bf16 weights do not stay in L2 from
        # one step to the next, and every kernel of this chain is latency bound at M = B*T
        pf = SideBranch(x.device, index=7) if (_L2_PREFETCH and dt == torch.bfloat16 and n_layers > 1) else None
        for l in range(n_layers):
            kqv_w, proj_w, g1, b1, w0, bb0, w2, bb2, g2, b2 = params[l * RF_PARAMS_PER_LAYER:(l + 1) * RF_PARAMS_PER_LAYER]
            wk = weight_cache.get((kqv_w,), dt)
            wp = weight_cache.get((proj_w,), dt)
            wf0 = weight_cache.get((w0,), dt)
            wf2 = weight_cache.get((w2,), dt)
            if pf is not None and l + 1 < n_layers:
                nx = params[(l + 1) * RF_PARAMS_PER_LAYER:(l + 2) * RF_PARAMS_PER_LAYER]
                with pf.after_now():
                    ops.l2_prefetch([weight_cache.get((nx[i],), dt) for i in (0, 1, 4, 6)], _L2_PREFETCH_CTAS)
            if fuse_kqv:
                attn, scores, kqv = ops.rf_attn_fwd_fused(xin, wk, scores, maskf, B, T, heads, d)
            else:
                kqv = torch.empty(M * heads, 3 * d, device=x.device, dtype=dt)
                ops.gemm(M * heads, 3 * d, d, xin, d, False, wk, d, False, kqv, 3 * d, b_static=True)
                attn, scores = ops.rf_attn_fwd(kqv, scores, maskf, B, T, heads, d)
            y1 = torch.empty(M, H, device=x.device, dtype=dt)
            ops.gemm(M, H, H, attn, H, False, wp, H, False, y1, H, epilogue=EPI_RESIDUAL, aux_in=xin, ld_aux_in=H,
                     dropout_p=p1, dropout_seed=seed + 2 * l, b_static=True)
            x1, _, mean1, rstd1 = ops.add_layernorm_fwd(y1, None, g1.detach(), b1.detach(), 1e-5, want_sum=False)
            F4 = wf0.shape[0]
            hpre = torch.empty(M, F4, device=x.device, dtype=dt)
            hact = torch.empty(M, F4, device=x.device, dtype=dt)
            ops.gemm(M, F4, H, x1, H, False, wf0, H, False, hact, F4, bias=bb0.detach(), epilogue=EPI_ACT, act=ACT_SERF,
                     aux_out=hpre, ld_aux_out=F4, b_static=True)
            ns = split_k_slabs(M, H, F4, x.device, dt)
            if ns > 1:
                # FF2 as split-K slabs: bias in slab 0; reduction + dropout + residual + LN2 in one pass
                if parts is None:
                    parts = torch.empty(ns, M, H, device=x.device, dtype=torch.float32)
                ops.gemm(M, H, F4, hact, F4, False, wf2, F4, False, parts, H, bias=bb2.detach(), split_k=ns,
                         c_split_stride=M * H, b_static=True)
                x2, y2, mean2, rstd2 = ops.add_layernorm_fwd_parts(parts, x1, g2.detach(), b2.detach(), 1e-5, dt, p2,
                                                                   seed + 2 * l + 1)
            else:
                y2 = torch.empty(M, H, device=x.device, dtype=dt)
                ops.gemm(M, H, F4, hact, F4, False, wf2, F4, False, y2, H, bias=bb2.detach(), epilogue=EPI_RESIDUAL, aux_in=x1,
                         ld_aux_in=H, dropout_p=p2, dropout_seed=seed + 2 * l + 1, b_static=True)
                x2, _, mean2, rstd2 = ops.add_layernorm_fwd(y2, None, g2.detach(), b2.detach(), 1e-5, want_sum=False)
            saved += [xin, kqv, scores, attn, y1, mean1, rstd1, x1, hpre, hact, y2, mean2, rstd2]
            xin = x2
        if pf is not None:
            pf.join()
        ctx.save_for_backward(*saved, *params)
        ctx.meta = (B, T, H, heads, d, n_layers, p1, p2, seed, dt, prev is not None)
        return xin.view(B, T, H), scores

    @staticmethod
    def backward(ctx, dy, dscores):
        B, T, H, heads, d, n_layers, p1, p2, seed, dt, has_prev = ctx.meta
        M = B * T
        nsave = 13
        saved = ctx.saved_tensors[:nsave * n_layers]
        params = ctx.saved_tensors[nsave * n_layers:]
        if dy is None:
            dx = torch.zeros(M, H, device=saved[0].device, dtype=dt)
        else:
            dx = dy.reshape(M, H).contiguous()
            if dx.dtype != dt:
                dx = dx.to(dt)
        ds = None if dscores is None else dscores.contiguous().float()
        grads: List[Optional[Tensor]] = [None] * len(params)
        # one zero-filled workspace for everything that is accumulated with atomics (LN gamma/beta, split-K kqv dW)
        F4 = params[4].shape[0]
        per_layer = 5 * H + F4 + 3 * d * d
        zws = torch.zeros(n_layers * per_layer, device=saved[0].device, dtype=torch.float32)
        # weight-gradient GEMMs are off the critical path (nothing in this backward reads them): they run on a side
        # branch and overlap the dgrad -> LayerNorm -> attention chain of the same and the following layers
        branch = SideBranch(saved[0].device)
        keep = []
        parts = None
        Tp_ = (T + 15) // 16 * 16
        fuse_kqv = (dt == torch.bfloat16 and d % 16 == 0 and T <= 128 and d <= 128 and
                    (2 * Tp_ * (d + 8) + 3 * d * (Tp_ + 8) + 2 * Tp_ * (Tp_ + 8) + Tp_ * (3 * d + 8) + 3 * d * (d + 8)) * 2
                    <= 200 * 1024 and _os.environ.get("MMVQA_FUSED_KQV_BWD") is not None)
        # (opt-in: measured at B = 16, T = 28 the fused backward is 4 us per layer SLOWER than attention backward + the
        # PDL-overlapped dgrad GEMM -- 120 KB of shared memory and a third serial phase on 128 CTAs; the forward fusion
        # is the one that pays)
        sink = _GRAD_SINK
        # opt-in (MMVQA_RF_ATTN_BWD=1): the attention block of every layer as one cluster kernel (csrc/rf_attn_block.cu).
        # Measured at B = 16, T = 28: it takes the same ~47 us as the four launches it replaces (one 8-CTA cluster per
        # sample pair = 64 SMs with 4 of 8 warps busy in the attention phases), 2.76 vs 2.71 ms/step -- parity-tested,
        # not the default.
        attn_block = (dt == torch.bfloat16 and _os.environ.get("MMVQA_RF_ATTN_BWD", "0") == "1" and
                      ops.rf_attn_block_bwd_supported(B, T, H, heads))
        pf = SideBranch(saved[0].device, index=7) if (_L2_PREFETCH and dt == torch.bfloat16 and n_layers > 1) else None
        for l in reversed(range(n_layers)):
            xin, kqv, scores, attn, y1, mean1, rstd1, x1, hpre, hact, y2, mean2, rstd2 = saved[l * nsave:(l + 1) * nsave]
            kqv_w, proj_w, g1, b1, w0, bb0, w2, bb2, g2, b2 = params[l * RF_PARAMS_PER_LAYER:(l + 1) * RF_PARAMS_PER_LAYER]
            wk = weight_cache.get((kqv_w,), dt)
            wp = weight_cache.get((proj_w,), dt)
            wf0 = weight_cache.get((w0,), dt)
            wf2 = weight_cache.get((w2,), dt)
            if pf is not None and l > 0:
                # the layer below: its weights and the activations its forward pass saved ~1 ms ago
                nx = params[(l - 1) * RF_PARAMS_PER_LAYER:l * RF_PARAMS_PER_LAYER]
                sv = saved[(l - 1) * nsave:l * nsave]
                with pf.after_now():
                    ops.l2_prefetch([weight_cache.get((nx[i],), dt) for i in (6, 4, 1, 0)] +
                                    [sv[i] for i in (10, 8, 9, 7, 4, 3, 1, 2, 0)], _L2_PREFETCH_CTAS)
            zl = zws[l * per_layer:(l + 1) * per_layer]
            dg2, db2, dg1, db1, dbb2 = zl[0:H], zl[H:2 * H], zl[2 * H:3 * H], zl[3 * H:4 * H], zl[4 * H:5 * H]
            dbb0 = zl[5 * H:5 * H + F4]
            # LN2 backward also emits dropout(dy2) for the FF branch and its column sums (= d ff.2.bias)
            # (column sums deferred: the kernel on the critical path stores per-CTA partials, the side branch folds them)
            dfr = ops.layernorm_bwd_deferred(dx, None, y2, g2.detach(), mean2, rstd2, want_drop=p2 > 0.0, dropout_p=p2,
                                             dropout_seed=seed + 2 * l + 1) if _LN_DEFER else None
            if dfr is not None:
                dy2, dff, lnp2 = dfr
                if dff is None:
                    dff = dy2
            elif p2 > 0.0:
                dy2, dff = ops.layernorm_bwd(dx, y2, g2.detach(), mean2, rstd2, None, dg2, db2, want_drop=True, dxsum=dbb2,
                                             dropout_p=p2, dropout_seed=seed + 2 * l + 1)
            else:
                dy2 = ops.layernorm_bwd(dx, y2, g2.detach(), mean2, rstd2, None, dg2, db2, dxsum=dbb2)
                dff = dy2
            # Opt-in (MMVQA_WGRAD_LATE=1): the two FF weight-gradient GEMMs enqueued AFTER the FF1 dgrad, so that they share
            # the chip with the LayerNorm / attention kernels instead of the FF dgrad GEMMs (FF2 dgrad takes 30 us inside
            # the step against 13 alone).  Measured neutral: the layer is bound by the sum of its serialized kernel
            # latencies, not by which kernels overlap.
            if not _WGRAD_LATE:
                with branch.after_now():
                    if dfr is not None:
                        ops.ln_partials_reduce(lnp2, dg2, db2, dbb2)
                        keep.append(lnp2)
                    dw2 = gemm_wgrad(dff, H, M, H, hact, F4, F4)
            # dgrad through FF2 with act'(h_pre) in the epilogue; its column sums are d ff.0.bias
            dhpre = gemm_dgrad(dff, H, M, H, wf2, F4, epilogue=EPI_DACT, act=ACT_SERF, aux_in=hpre, colsum_out=dbb0)
            if not _WGRAD_LATE:
                with branch.after_now():
                    dw0 = gemm_wgrad(dhpre, F4, M, F4, x1, H, H)
            ns = split_k_slabs(M, H, F4, dx.device, dt)
            if attn_block:
                # LN1 backward + proj dgrad + attention backward + kqv dgrad in ONE cluster launch (csrc/rf_attn_block.cu):
                # it consumes the fp32 split-K slabs of the FF1 dgrad and dy2 directly
                if parts is None:
                    parts = torch.empty(max(ns, 1), M, H, device=dx.device, dtype=torch.float32)
                ops.gemm(M, H, F4, dhpre, F4, False, wf0, H, True, parts, H, split_k=max(ns, 1), c_split_stride=M * H,
                         b_static=True)
                if _WGRAD_LATE:
                    with branch.after_now():
                        if dfr is not None:
                            ops.ln_partials_reduce(lnp2, dg2, db2, dbb2)
                            keep.append(lnp2)
                        dw2 = gemm_wgrad(dff, H, M, H, hact, F4, F4)
                        dw0 = gemm_wgrad(dhpre, F4, M, F4, x1, H, H)
                want_dprev = (l > 0) or (has_prev and ctx.needs_input_grad[2])
                dpr, dkqv, dprev, dxin = ops.rf_attn_block_bwd(parts, dy2, y1, mean1, rstd1, g1.detach(), wp, wk, kqv, scores, ds,
                                                               want_dprev, dg1, db1, B, T, heads, p1, seed + 2 * l)
                with branch.after_now():
                    dwp = gemm_wgrad(dpr, H, M, H, attn, H, H)
                    dwk = gemm_wgrad(dkqv, 3 * d, M * heads, 3 * d, xin, d, d, zeroed=zl[5 * H + F4:].view(3 * d, d))
                keep.append((dff, dhpre, dpr, dkqv, dy2))
                base = l * RF_PARAMS_PER_LAYER
                gl = [dwk.view(kqv_w.shape), dwp.view(proj_w.shape), dg1, db1, dw0.view(w0.shape), dbb0, dw2.view(w2.shape), dbb2,
                      dg2, db2]
                pl = params[base:base + RF_PARAMS_PER_LAYER]
                if sink is not None and all(ctx.needs_input_grad[7 + base + i] for i in range(RF_PARAMS_PER_LAYER)) and \
                        sink(pl, gl, branch.side, l == 0):
                    keep.append(gl)
                else:
                    grads[base:base + RF_PARAMS_PER_LAYER] = gl
                dx = dxin
                ds = dprev
                continue
            if ns > 1:
                # dgrad through FF1 as split-K slabs; LN1 backward sums them and adds the residual gradient dy2
                if parts is None:
                    parts = torch.empty(ns, M, H, device=dx.device, dtype=torch.float32)
                ops.gemm(M, H, F4, dhpre, F4, False, wf0, H, True, parts, H, split_k=ns, c_split_stride=M * H, b_static=True)
                if _WGRAD_LATE:
                    with branch.after_now():
                        if dfr is not None:
                            ops.ln_partials_reduce(lnp2, dg2, db2, dbb2)
                            keep.append(lnp2)
                        dw2 = gemm_wgrad(dff, H, M, H, hact, F4, F4)
                        dw0 = gemm_wgrad(dhpre, F4, M, F4, x1, H, H)
                dfr1 = ops.layernorm_bwd_deferred(dy2, parts, y1, g1.detach(), mean1, rstd1, want_drop=p1 > 0.0, dropout_p=p1,
                                                  dropout_seed=seed + 2 * l) if _LN_DEFER else None
                if dfr1 is not None:
                    dy1, dpr, lnp1 = dfr1
                    if dpr is None:
                        dpr = dy1
                    with branch.after_now():
                        ops.ln_partials_reduce(lnp1, dg1, db1, None)
                    keep.append(lnp1)
                elif p1 > 0.0:
                    dy1, dpr = ops.layernorm_bwd_parts(parts, dy2, y1, g1.detach(), mean1, rstd1, dg1, db1, want_drop=True,
                                                       dropout_p=p1, dropout_seed=seed + 2 * l)
                else:
                    dy1 = ops.layernorm_bwd_parts(parts, dy2, y1, g1.detach(), mean1, rstd1, dg1, db1)
                    dpr = dy1
            else:
                dx1 = gemm_dgrad(dhpre, F4, M, F4, wf0, H, epilogue=EPI_RESIDUAL, aux_in=dy2)
                if _WGRAD_LATE:
                    with branch.after_now():
                        if dfr is not None:
                            ops.ln_partials_reduce(lnp2, dg2, db2, dbb2)
                            keep.append(lnp2)
                        dw2 = gemm_wgrad(dff, H, M, H, hact, F4, F4)
                        dw0 = gemm_wgrad(dhpre, F4, M, F4, x1, H, H)
                dfr1 = ops.layernorm_bwd_deferred(dx1, None, y1, g1.detach(), mean1, rstd1, want_drop=p1 > 0.0, dropout_p=p1,
                                                  dropout_seed=seed + 2 * l) if _LN_DEFER else None
                if dfr1 is not None:
                    dy1, dpr, lnp1 = dfr1
                    if dpr is None:
                        dpr = dy1
                    with branch.after_now():
                        ops.ln_partials_reduce(lnp1, dg1, db1, None)
                    keep.append(lnp1)
                elif p1 > 0.0:
                    dy1, dpr = ops.layernorm_bwd(dx1, y1, g1.detach(), mean1, rstd1, None, dg1, db1, want_drop=True,
                                                 dropout_p=p1, dropout_seed=seed + 2 * l)
                else:
                    dy1 = ops.layernorm_bwd(dx1, y1, g1.detach(), mean1, rstd1, None, dg1, db1)
                    dpr = dy1
            with branch.after_now():
                dwp = gemm_wgrad(dpr, H, M, H, attn, H, H)
            dattn = gemm_dgrad(dpr, H, M, H, wp, H)
            want_dprev = (l > 0) or (has_prev and ctx.needs_input_grad[2])
            if fuse_kqv:
                # dx_in = dkqv . Wkqv (per head) + dy1 comes out of the attention backward kernel itself
                dkqv, dprev, dxin = ops.rf_attn_bwd_fused(kqv, scores, dattn, ds, want_dprev, wk, dy1, B, T, heads, d)
            else:
                dkqv, dprev = ops.rf_attn_bwd(kqv, scores, dattn, ds, want_dprev, B, T, heads, d)
            with branch.after_now():
                dwk = gemm_wgrad(dkqv, 3 * d, M * heads, 3 * d, xin, d, d, zeroed=zl[5 * H + F4:].view(3 * d, d))
            keep.append((dff, dhpre, dpr, dkqv, dy2, dy1))
            if not fuse_kqv:
                # dx_in = dkqv . Wkqv (per head) + dy1 (residual around the attention block)
                dxin = torch.empty(M, H, device=dx.device, dtype=dt)
                ops.gemm(M * heads, d, 3 * d, dkqv, 3 * d, False, wk, d, True, dxin, d, epilogue=EPI_RESIDUAL, aux_in=dy1,
                         ld_aux_in=d, b_static=True)
            base = l * RF_PARAMS_PER_LAYER
            gl = [dwk.view(kqv_w.shape), dwp.view(proj_w.shape), dg1, db1, dw0.view(w0.shape), dbb0, dw2.view(w2.shape), dbb2,
                  dg2, db2]
            pl = params[base:base + RF_PARAMS_PER_LAYER]
            if sink is not None and all(ctx.needs_input_grad[7 + base + i] for i in range(RF_PARAMS_PER_LAYER)) and \
                    sink(pl, gl, branch.side, l == 0):
                keep.append(gl)
            else:
                grads[base:base + RF_PARAMS_PER_LAYER] = gl
            dx = dxin
            ds = dprev
        branch.join()
        if pf is not None:
            pf.join()
        del keep
        dprev_out = ds if (has_prev and ctx.needs_input_grad[2]) else None
        return (dx.view(B, T, H), None, dprev_out, None, None, None, None, *grads)


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
class ASLFn(torch.autograd.Function):
    """ASLSingleLabel (models/asl_singlelabel.py:23-52): per-row loss + its gradient in one launch."""

    @staticmethod
    def forward(ctx, logits: Tensor, target: Tensor, gp: float, gn: float, eps: float, want_tc: bool):
        lg = logits if logits.is_contiguous() else logits.contiguous()
        Bn, Cn = lg.shape
        loss_rows, dl, tc = ops.asl_fwd_bwd(lg, Cn, target.contiguous().long(), Cn, gp, gn, eps, True, want_tc)
        ctx.save_for_backward(dl)
        ctx.in_dtype = logits.dtype
        if tc is None:
            tc = torch.empty(0, device=lg.device)
        ctx.mark_non_differentiable(tc)
        return loss_rows, tc

    @staticmethod
    def backward(ctx, dloss_rows, _dtc):
        (dl,) = ctx.saved_tensors
        g = dl * dloss_rows.reshape(-1, 1).float()
        return g.to(ctx.in_dtype), None, None, None, None, None


class CrossEntropyRowsFn(torch.autograd.Function):
    """NLL(log_softmax(logits)) per row (pretrain/roco_utils.py:235-236) with the gradient produced in
    the same pass; `ld` lets the logits live in a padded buffer."""

    @staticmethod
    def forward(ctx, logits2d: Tensor, target: Tensor):
        rows, Cn = logits2d.shape
        ld = logits2d.stride(0)
        if logits2d.stride(1) != 1:
            logits2d = logits2d.contiguous()
            ld = Cn
        dl = torch.empty(rows, ld, device=logits2d.device, dtype=logits2d.dtype)
        loss_rows = ops.ce_fwd_bwd(logits2d, ld, target.contiguous().long(), rows, Cn, 1.0, dl, ld)
        ctx.save_for_backward(dl)
        ctx.Cn = Cn
        return loss_rows

    @staticmethod
    def backward(ctx, dloss_rows):
        (dl,) = ctx.saved_tensors
        g = dl[:, :ctx.Cn] * dloss_rows.reshape(-1, 1).to(dl.dtype)
        return g, None


class ChunkedVocabCEFn(torch.autograd.Function):
    """Per-row NLL of log_softmax(h W^T + b) (mmbert.py:154-155 + pretrain/roco_utils.py:235-236) WITHOUT the [M, V] logits:
    the vocabulary is walked in column chunks -- vocab GEMM into an [M, chunk] fp32 scratch, online log-sum-exp
    (mmvqa_ce_chunk_stats); the backward pass recomputes each chunk, turns it into dlogits (mmvqa_ce_chunk_grad) and
    feeds it straight to the dgrad (dh += dl W_c) and wgrad (dW_c = dl^T h, db_c = colsum dl) GEMMs.  Memory is
    M * chunk * 6 bytes instead of two [M, V] fp32 tensors (2 x 293 MB at M = 2400); the price is one extra vocab GEMM."""

    @staticmethod
    def forward(ctx, h: Tensor, weight: Tensor, bias: Tensor, target: Tensor, chunk: int):
        dt = h.dtype
        M, H = h.shape
        V = weight.shape[0]
        hc = h if h.is_contiguous() else h.contiguous()
        w = weight_cache.get((weight,), dt)
        tgt = target.contiguous().long()
        scratch = torch.empty(M, chunk, device=h.device, dtype=torch.float32)
        rowmax = torch.empty(M, device=h.device, dtype=torch.float32)
        rowsum = torch.empty(M, device=h.device, dtype=torch.float32)
        tl = torch.zeros(M, device=h.device, dtype=torch.float32)
        b32 = bias.detach().float()
        for c0 in range(0, V, chunk):
            vc = min(chunk, V - c0)
            ops.gemm(M, vc, H, hc, H, False, w[c0:c0 + vc], H, False, scratch, chunk, bias=b32[c0:c0 + vc], b_static=True)
            ops.ce_chunk_stats(scratch, chunk, tgt, M, c0, vc, rowmax, rowsum, tl, c0 == 0)
        ctx.save_for_backward(hc, weight, bias, tgt, rowmax, rowsum)
        ctx.meta = (chunk, dt)
        return rowmax + torch.log(rowsum) - tl

    @staticmethod
    def backward(ctx, dloss_rows):
        hc, weight, bias, tgt, rowmax, rowsum = ctx.saved_tensors
        chunk, dt = ctx.meta
        M, H = hc.shape
        V = weight.shape[0]
        w = weight_cache.get((weight,), dt)
        b32 = bias.detach().float()
        rs = dloss_rows.reshape(-1).contiguous().float()
        scratch = torch.empty(M, chunk, device=hc.device, dtype=torch.float32)
        dl = torch.empty(M, chunk, device=hc.device, dtype=dt)
        dh = torch.zeros(M, H, device=hc.device, dtype=torch.float32)
        dW = torch.empty(V, H, device=hc.device, dtype=torch.float32)
        db = torch.empty(V, device=hc.device, dtype=torch.float32)
        for c0 in range(0, V, chunk):
            vc = min(chunk, V - c0)
            ops.gemm(M, vc, H, hc, H, False, w[c0:c0 + vc], H, False, scratch, chunk, bias=b32[c0:c0 + vc], b_static=True)
            ops.ce_chunk_grad(scratch, chunk, tgt, M, c0, vc, rowmax, rowsum, rs, dl, chunk)
            ops.gemm(M, H, vc, dl, chunk, False, w[c0:c0 + vc], H, True, dh, H, accumulate=True, b_static=True)
            ops.gemm(vc, H, M, dl, chunk, True, hc, H, True, dW[c0:c0 + vc], H)
            db[c0:c0 + vc] = ops.colsum(dl, M, vc, chunk)
        return dh.to(dt), dW, db, None, None


def chunked_vocab_ce(h: Tensor, weight: Tensor, bias: Tensor, target: Tensor, chunk: int = 4096) -> Tensor:
    """per-row MLM losses [M] for hidden states h [M, H] (compute dtype) and the vocabulary projection (weight [V, H],
    bias [V]); see ChunkedVocabCEFn."""
    return ChunkedVocabCEFn.apply(to_compute(h), weight, bias, target, int(chunk))


class SupConFn(torch.autograd.Function):
    """SupConLoss core (models/SupConLoss/loss.py:58-96) on the view-major contrast matrix F [N, D]:
    raw = anchors . F^T on the tensor cores, then one fused row pass (max, masked exp-sum, positives
    mean, gradient).  Returns the per-anchor losses."""

    @staticmethod
    def forward(ctx, Fm: Tensor, mask: Optional[Tensor], bsz: int, n_anchor_rows: int, temperature: float,
                base_temperature: float, dtype: torch.dtype, row_offset: int = 0):
        """anchors = rows [row_offset, row_offset + n_anchor_rows) of the contrast matrix (row_offset > 0: the
        local-anchor slice of a data-parallel rank, mmvqa_b200.parallel.supcon_loss_sharded)."""
        N, D = Fm.shape
        Fc, ld = _pad_ld(Fm.detach(), dtype)
        raw = torch.empty(n_anchor_rows, N, device=Fm.device, dtype=torch.float32)
        ops.gemm(n_anchor_rows, N, D, Fc[row_offset:], ld, False, Fc, ld, False, raw, N)
        loss_rows, G = ops.supcon_rows(raw, None if mask is None else mask.contiguous().float(), bsz, row_offset, temperature,
                                       base_temperature, True)
        ctx.save_for_backward(Fc, G)
        ctx.meta = (N, D, ld, n_anchor_rows, dtype, Fm.dtype, row_offset)
        return loss_rows

    @staticmethod
    def backward(ctx, dloss_rows):
        Fc, G = ctx.saved_tensors
        N, D, ld, R, dtype, in_dtype, off = ctx.meta
        Gs = G * dloss_rows.reshape(-1, 1).float()
        Gc, ldg = _pad_ld(Gs, dtype)
        dF = torch.zeros(N, D, device=Fc.device, dtype=torch.float32)
        # anchor role: dF[off:off+R] += G . F ;  contrast role: dF += G^T . F[off:off+R]
        ops.gemm(R, D, N, Gc, ldg, False, Fc, ld, True, dF[off:], D, accumulate=True)
        ops.gemm(N, D, R, Gc, ldg, True, Fc[off:], ld, True, dF, D, accumulate=True)
        return dF.to(in_dtype), None, None, None, None, None, None, None
