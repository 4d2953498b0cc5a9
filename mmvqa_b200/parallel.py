"""Batch-sharded data parallelism: one process per GPU, full weight replica, two exchanges
per step -- a bucketed gradient all-reduce and (SupCon only) an all-gather of the [N_local, D]
embeddings with the matching gradient reduction.  torch.distributed (NCCL over NVLink 5 / NVSwitch on the
GPU box, gloo in the CPU tests) is the transport; the reference has no distributed code at all."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """identical init on every rank (rank-0 state wins)."""
    if not is_dist():
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)
    # writes through .data do not bump Parameter._version: the cached bf16 operand copies of the non-source ranks
    # would silently keep the pre-broadcast weights
    from .functional import invalidate_weight_cache
    invalidate_weight_cache()


class GradBuckets:
    """Flat fp32 gradient buckets.  After backward: ``reduce()`` packs every p.grad into its bucket (in reverse
    parameter order, so the first bucket holds the gradients that were produced last and every earlier
    all-reduce overlaps the packing of the next bucket), launches one async all-reduce(SUM) per bucket and
    waits.  ``grads()`` returns per-parameter views into the buckets for FusedAdam(step(grads=...)) with
    ``grad_scale = 1/world`` -- gradients are never copied back."""

    def __init__(self, params, bucket_mb: float = 64.0, dtype: torch.dtype = torch.float32):
        self.params = [p for p in params if p.requires_grad]
        cap = int(bucket_mb * 1024 * 1024 / (2 if dtype == torch.bfloat16 else 4))
        self.buckets: List[torch.Tensor] = []
        self.slots = []                      # (param index, bucket index, offset, numel)
        cur, cur_n = [], 0
        order = list(reversed(range(len(self.params))))
        groups = []
        for i in order:
            n = (self.params[i].numel() + 7) // 8 * 8          # keep 16-byte alignment of every view (fp32 and bf16)
            if cur and cur_n + n > cap:
                groups.append((cur, cur_n))
                cur, cur_n = [], 0
            cur.append((i, cur_n, self.params[i].numel()))
            cur_n += n
        if cur:
            groups.append((cur, cur_n))
        dev = self.params[0].device
        for bi, (items, total) in enumerate(groups):
            self.buckets.append(torch.zeros(total, device=dev, dtype=dtype))
            for (i, off, n) in items:
                self.slots.append((i, bi, off, n))
        self._views = {}
        for (i, bi, off, n) in self.slots:
            self._views[i] = self.buckets[bi][off:off + n].view_as(self.params[i])

    def pack(self) -> None:
        """copy every p.grad into its bucket slot (pure device copies: CUDA-graph capturable)."""
        per_bucket = {}
        for (i, bi, off, n) in self.slots:
            per_bucket.setdefault(bi, []).append(i)
        for bi, idxs in per_bucket.items():
            dsts = [self._views[i] for i in idxs]
            srcs = [self.params[i].grad if self.params[i].grad is not None else torch.zeros_like(self.params[i]) for i in idxs]
            torch._foreach_copy_(dsts, srcs)

    def allreduce(self) -> None:
        """one async all-reduce(SUM) per bucket, then wait (NCCL over NVLink on the GPU box, gloo in CPU tests)."""
        if not is_dist():
            return
        works = [dist.all_reduce(b, op=dist.ReduceOp.SUM, async_op=True) for b in self.buckets]
        for w in works:
            w.wait()

    def reduce(self) -> None:
        self.pack()
        self.allreduce()

    def grads(self, plist=None) -> List[torch.Tensor]:
        idx = {id(p): i for i, p in enumerate(self.params)}
        plist = self.params if plist is None else plist
        return [self._views[idx[id(p)]] for p in plist]

    @property
    def grad_scale(self) -> float:
        return 1.0 / (dist.get_world_size() if is_dist() else 1)


class LayerwiseReducer:
    """Gradient exchange at layer granularity, for ``FusedAdam(reduce_fn=...)``.

    ``reducer(params, grads)`` packs the gradients of one set of parameters (one encoder layer when it is called
    from the backward gradient sink, everything else at ``step()``) into that set's persistent flat bucket
    (fp32 -> bf16 cast fused into the pack when ``dtype`` is bfloat16), all-reduces the bucket (SUM; the caller
    sets ``FusedAdam.grad_scale = 1 / world``) and returns per-parameter views of it.  Called on the stream the
    update runs on, so with ``overlap_backward`` the all-reduce of layer l travels over NVLink while layers
    l-1 ... 0 are still in backward; the whole data-parallel step (NCCL kernels included) is one CUDA graph."""

    def __init__(self, dtype: torch.dtype = torch.float32, multimem: bool = False, multimem_ctas: int = 16):
        self.dtype = dtype
        self._buckets = {}
        self.bytes_per_step = 0
        # multimem: buckets in symmetric memory, all-reduced in the NVSwitch by the library's own kernel
        # (mmvqa_multimem_allreduce) instead of NCCL ring kernels
        self.multimem = bool(multimem) and is_dist() and dist.get_backend() == "nccl"
        self.multimem_ctas = multimem_ctas
        self._handles = {}
        self._gathered = {}        # id(param) -> all ranks' ids of the last exchange
        self._last_ids = {}        # id(param) -> persistent copy of those ids (rows to clear before the next exchange)
        self._row_sparse = {}      # id(param) -> callable returning the int64 row ids this rank touched in the step
        self._dense = {}           # id(param) -> persistent dense gradient buffer of a row-sparse parameter

    def register_row_sparse(self, param: torch.nn.Parameter, ids_fn) -> None:
        """`param` is an embedding table [V, H] whose gradient is zero outside the rows ``ids_fn()`` (any shape, int64,
        duplicates allowed, the SAME number of ids on every rank; a fixed device tensor when the step is graph-captured):
        BertEmbeddings.word_embeddings with the batch's input_ids (models/mmbert.py:52-63).  Instead of all-reducing the
        dense table gradient (47 MB in bf16 for bert-base, exchanged at the very end of the backward pass, fully exposed)
        the ranks all-gather the touched rows (<= B*T rows of H values each) and their ids and scatter-add them into a
        local dense buffer: the result equals the dense all-reduce(SUM) and Adam reads it exactly as before."""
        self._row_sparse[id(param)] = ids_fn

    def gathered_ids(self, param) -> torch.Tensor:
        """ids of every rank from the last row exchange of `param` (what FusedAdam.register_row_sparse needs under
        data parallelism: the rows that hold gradient after the exchange)."""
        return self._gathered[id(param)]

    def _exchange_rows(self, param, grad):
        ids = self._row_sparse[id(param)]().reshape(-1)
        n = ids.numel()
        # one representative per distinct id (the dense local gradient row already holds the sum over its duplicates)
        first = (ids.unsqueeze(0) == ids.unsqueeze(1)).to(torch.int32).argmax(dim=1)
        keep = (first == torch.arange(n, device=ids.device)).to(self.dtype).unsqueeze(1)
        rows = grad.index_select(0, ids).to(self.dtype) * keep
        dense = self._dense.get(id(param))
        if dense is None:
            # fp32 accumulator: rows shared by every rank ([CLS], [SEP]) are summed world times
            dense = self._dense[id(param)] = torch.zeros(grad.shape, device=grad.device, dtype=torch.float32)
            self.bytes_per_step += (rows.numel() * rows.element_size() + ids.numel() * 8)
        world = dist.get_world_size() if is_dist() else 1
        if world > 1:
            all_rows = torch.empty((world * n, rows.shape[1]), device=rows.device, dtype=rows.dtype)
            all_ids = torch.empty(world * n, device=ids.device, dtype=ids.dtype)
            if dist.get_backend() == "nccl":
                dist.all_gather_into_tensor(all_rows, rows.contiguous())
                dist.all_gather_into_tensor(all_ids, ids.contiguous())
            else:
                dist.all_gather(list(all_rows.chunk(world, dim=0)), rows.contiguous())
                dist.all_gather(list(all_ids.chunk(world, dim=0)), ids.contiguous())
        else:
            all_rows, all_ids = rows, ids
        # the dense accumulator is all zeros except the rows of the previous exchange: clear those (<= world * n rows)
        # instead of the whole 94 MB table.  Their ids live in a persistent buffer so that a captured step clears what its
        # previous REPLAY wrote, not what the capture-time warm-up wrote.
        last = self._last_ids.get(id(param))
        if last is None or last.numel() != all_ids.numel():
            last = self._last_ids[id(param)] = torch.empty_like(all_ids)
            dense.zero_()
        else:
            dense.index_fill_(0, last, 0.0)
        last.copy_(all_ids)
        self._gathered[id(param)] = last
        dense.index_add_(0, all_ids, all_rows.float())
        return dense

    def __call__(self, params, grads):
        sparse = [i for i, p in enumerate(params) if id(p) in self._row_sparse]
        if sparse:
            dense_idx = [i for i in range(len(params)) if i not in sparse]
            out = [None] * len(params)
            if dense_idx:
                for i, v in zip(dense_idx, self._reduce_dense([params[i] for i in dense_idx], [grads[i] for i in dense_idx])):
                    out[i] = v
            for i in sparse:
                out[i] = self._exchange_rows(params[i], grads[i])
            return out
        return self._reduce_dense(params, grads)

    def _reduce_dense(self, params, grads):
        key = (id(params[0]), len(params))
        ent = self._buckets.get(key)
        if ent is None:
            offs, total = [], 0
            for g in grads:
                offs.append(total)
                total += (g.numel() + 7) // 8 * 8          # 16-byte aligned views for fp32 and bf16
            flat = None
            if self.multimem:
                import torch.distributed._symmetric_memory as symm_mem
                flat = symm_mem.empty(total, dtype=self.dtype, device=grads[0].device)
                hdl = symm_mem.rendezvous(flat, dist.group.WORLD)          # collective: same bucket order on every rank
                if getattr(hdl, "multicast_ptr", 0):
                    flat.zero_()
                    self._handles[key] = hdl
                else:
                    flat = None                                            # no multicast support: NCCL
            if flat is None:
                flat = torch.zeros(total, device=grads[0].device, dtype=self.dtype)
            views = [flat[o:o + g.numel()].view_as(g) for o, g in zip(offs, grads)]
            ent = self._buckets[key] = (flat, views)
            self.bytes_per_step += flat.numel() * flat.element_size()
        flat, views = ent
        torch._foreach_copy_(views, list(grads))
        if is_dist():
            hdl = self._handles.get(key)
            if hdl is not None:
                from . import ops
                ops.multimem_allreduce(hdl.multicast_ptr, hdl.signal_pad_ptrs_dev, hdl.rank, hdl.world_size,
                                       flat.numel() * flat.element_size(), self.dtype, self.multimem_ctas)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return views

    @property
    def grad_scale(self) -> float:
        return 1.0 / (dist.get_world_size() if is_dist() else 1)


class _GatherFeatures(torch.autograd.Function):
    """all-gather of per-rank SupCon embeddings [n_local, n_views, D] -> [n_global, n_views, D] (rank-major, which
    keeps the reference's view-major contrast order inside SupConLoss: all view-0 rows, then all view-1 rows).
    Backward: sum the contrast-role gradients of every rank (all-reduce) and keep the local slice."""

    @staticmethod
    def forward(ctx, feat):
        world, rank = dist.get_world_size(), dist.get_rank()
        outs = [torch.empty_like(feat) for _ in range(world)]
        dist.all_gather(outs, feat.contiguous())
        ctx.meta = (rank, feat.shape[0])
        return torch.cat(outs, dim=0)

    @staticmethod
    def backward(ctx, dgathered):
        rank, n = ctx.meta
        g = dgathered.contiguous()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return g[rank * n:(rank + 1) * n]


def gather_supcon_features(feat: torch.Tensor) -> torch.Tensor:
    """feat [n_local, n_views, D] -> global [world * n_local, n_views, D].  The SupCon loss of the gathered tensor
    is the same on every rank; with DP gradient averaging (sum / world) each rank back-propagates
    world * loss / world = the global-batch loss gradient through its local slice."""
    if not is_dist():
        return feat
    return _GatherFeatures.apply(feat)


def gather_mask_rows(mask_local_rows: torch.Tensor) -> torch.Tensor:
    """rows of the [bsz_global, bsz_global] similarity mask owned by this rank -> full mask (no gradient)."""
    if not is_dist():
        return mask_local_rows
    outs = [torch.empty_like(mask_local_rows) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, mask_local_rows.contiguous())
    return torch.cat(outs, dim=0)


class _GatherFeaturesRS(torch.autograd.Function):
    """all-gather of [n_local, n_views, D] embeddings (rank-major) whose backward is a REDUCE-SCATTER (sum): with
    local-anchor rows every rank holds a different gradient of the gathered tensor -- the anchor-role rows of its own
    samples and the contrast-role contributions to everybody else's -- and each rank needs only the sum over ranks of
    its own slice (collective 2)."""

    @staticmethod
    def forward(ctx, feat):
        world = dist.get_world_size()
        feat = feat.contiguous()
        out = torch.empty((world * feat.shape[0],) + tuple(feat.shape[1:]), device=feat.device, dtype=feat.dtype)
        if dist.get_backend() == "nccl":
            dist.all_gather_into_tensor(out, feat)
        else:
            dist.all_gather(list(out.chunk(world, dim=0)), feat)
        ctx.n = feat.shape[0]
        return out

    @staticmethod
    def backward(ctx, dgathered):
        n, rank = ctx.n, dist.get_rank()
        g = dgathered.contiguous()
        if dist.get_backend() == "nccl":
            out = torch.empty((n,) + tuple(g.shape[1:]), device=g.device, dtype=g.dtype)
            dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM)
            return out
        dist.all_reduce(g, op=dist.ReduceOp.SUM)          # gloo has no reduce-scatter: all-reduce + slice
        return g[rank * n:(rank + 1) * n]


def supcon_loss_sharded(feat_local: torch.Tensor, mask: Optional[torch.Tensor] = None, temperature: float = 0.07,
                        base_temperature: float = 0.07) -> torch.Tensor:
    """SupConLoss (models/SupConLoss/loss.py:21-98, contrast_mode 'all') of the GLOBAL batch with the anchor rows
    partitioned over the ranks: every rank all-gathers the embeddings, computes the similarity rows of ITS samples'
    anchors against all world * n_local * n_views contrasts (1 / world of the N x N problem) and returns

        loss_r = world * sum_{local anchors} row_loss / (n_views * bsz_global),

    so that mean_r loss_r is the reference's loss of the global batch (loss.py:96) and, with the usual data-parallel
    gradient averaging, every replica receives the global-batch gradient: the backward of the gather is a reduce-scatter
    of the contrast-role gradients.  `mask`: optional FULL [bsz_global, bsz_global] float mask (gather_mask_rows).
    Single process: identical to SupConLoss()(feat_local, mask=mask)."""
    from . import functional as Fn
    from .config import compute_dtype
    if feat_local.dim() < 3:
        raise ValueError('`features` needs to be [bsz, n_views, ...],at least 3 dimensions are required')
    feat_local = feat_local.reshape(feat_local.shape[0], feat_local.shape[1], -1)
    world = dist.get_world_size() if is_dist() else 1
    rank = dist.get_rank() if is_dist() else 0
    n, nv = feat_local.shape[0], feat_local.shape[1]
    gathered = _GatherFeaturesRS.apply(feat_local) if world > 1 else feat_local
    bsz = world * n
    if mask is not None and tuple(mask.shape) != (bsz, bsz):
        raise ValueError("supcon_loss_sharded: mask must be the full [bsz_global, bsz_global] matrix")
    Fm = gathered.transpose(0, 1).reshape(nv * bsz, -1)                 # view-major contrast matrix (loss.py:58)
    if Fm.dtype not in (torch.float32, torch.bfloat16):
        Fm = Fm.float()
    Fm = Fm.contiguous()
    m = None if mask is None else mask.float().to(Fm.device)
    total = None
    for v in range(nv):
        rows = Fn.SupConFn.apply(Fm, m, bsz, n, float(temperature), float(base_temperature), compute_dtype(),
                                 v * bsz + rank * n)
        total = rows.sum() if total is None else total + rows.sum()
    return total * (float(world) / float(nv * bsz))
