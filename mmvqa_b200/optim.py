"""Fused multi-tensor Adam: torch.optim.Adam semantics (vqamed2019/train.py:160) in ONE
kernel launch over every parameter, with a state_dict compatible with torch.optim.Adam so the
reference's ``recorder_2.pt`` checkpoints round-trip (pretrain/roco_train.py:164-171)."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import functional as Fn
from . import ops
from .functional import weight_cache

CHUNK = 8192             # elements per table entry of an update on the critical path (two 16-byte groups in flight)
CHUNK_BACKGROUND = 32768  # ... of an update that runs underneath the backward pass (one group in flight, long gentle CTAs)
SMALL_UPDATE = 1 << 22   # parameters: below this an update launch uses the critical-path mode
GATED_ROWS = 256         # rows per entry of a row-gated embedding table (a warp scans 32 row flags per load)
_DESC = np.dtype([("p", "<u8"), ("m", "<u8"), ("v", "<u8"), ("g", "<u8"), ("bf16_out", "<u8"), ("n", "<i8"), ("flags", "<i8"),
                  ("row_live", "<u8"), ("row_len", "<i8")])


class FusedAdam(torch.optim.Optimizer):
    """Drop-in for ``torch.optim.Adam(params, lr)`` (no amsgrad / maximize / capturable flags).

    ``grad_scale`` multiplies every gradient inside the kernel (1/world_size after an all-reduce SUM).
    The step counter, every group's ``lr`` and ``grad_scale`` live on the device, so a captured CUDA graph of
    ``step()`` can be replayed and still follows ``ReduceLROnPlateau`` / manual ``param_groups[i]['lr']`` changes
    (vqamed2019/train.py:161,233): ``refresh_hyper()`` re-uploads them when they changed -- eager ``step()`` calls it
    itself, ``GraphedTrainStep.replay`` calls it before every replay.  betas / eps / weight_decay are kernel
    arguments: changing them after a capture needs a re-capture."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, overlap_backward=False,
                 reduce_fn=None, early_groups=None, sink_group: int = 1):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = 1.0
        self._tables = {}
        self._row_gate = {}        # id(param) -> {"live": uint8 [rows], "ids_fn": callable}  (register_row_sparse)
        self._keepalive = []       # pinned host tables referenced by memcpy nodes of captured graphs
        self._eager_tables = {}    # last eagerly built table per key (kept until it is rebuilt)
        self._refreshed = {}       # group index -> ids of parameters whose bf16 cache copy the kernel rewrites
        self._step_dev = None
        self._hyper_dev = None     # [n_groups, 2] fp32 on the device: lr, grad_scale
        self._hyper_host = None    # pinned mirror
        # overlap_backward: layers whose backward node offers its gradients early (functional.set_grad_sink) are
        # updated on a separate stream while the rest of the backward pass runs -- Adam is pure HBM traffic, the
        # small-batch backward is latency bound, so the two overlap almost for free.  step() updates what is left.
        # reduce_fn(params, grads) -> grads' (data parallel): called right before the update of a set of parameters,
        # on the stream that update runs on; packs / all-reduces and returns the tensors Adam should read
        # (mmvqa_b200.parallel.LayerwiseReducer).  With overlap_backward each layer's exchange overlaps the backward.
        self.reduce_fn = reduce_fn
        # grid cap of the updates that run underneath the backward pass (0 = one CTA per CHUNK elements)
        self.early_ctas = int(os.environ.get("MMVQA_ADAM_EARLY_CTAS", "0"))
        self._stepped = False      # device step counter already advanced in this iteration
        self._early_ids = set()    # parameters already updated in this iteration
        self._early_keep = []      # gradients read by the optimizer stream (kept alive until step() joins it)
        self._opt_stream = None
        self._hook_stream = None
        self._group_of = {id(p): gi for gi, g in enumerate(self.param_groups) for p in g["params"]}
        self._comm_stream = None
        self._hook_groups = []
        # sink_group: how many consecutive gradient-sink calls (layers) share one exchange + one update launch.  With a
        # reduce_fn, fewer and larger all-reduces use the links better and occupy the SMs for less time in total.
        # A sequence is a schedule: the i-th exchange takes sink_group[i] layers (the last entry repeats), e.g. (4, 4, 2, 1, 1)
        # -- large exchanges while many layers of backward remain to hide them, single layers at the end of the backward
        # pass, where the last exchange + update is exposed.
        if isinstance(sink_group, (list, tuple)):
            self._sink_schedule = [max(1, int(x)) for x in sink_group] or [1]
        else:
            self._sink_schedule = [max(1, int(sink_group))]
        self.sink_group = self._sink_schedule[0]
        self._sink_round = 0
        self._pending = []         # [(params, grads)] handed in by the sink, not launched yet
        if overlap_backward:
            Fn.set_grad_sink(self._sink)
            # early_groups: lists of ordinary parameters (e.g. the BertEmbeddings tables, whose gradients appear before
            # the projector backward starts) that are updated as soon as autograd has accumulated the whole list
            for grp in (early_groups or []):
                grp = [p for p in grp if p.requires_grad]
                if not grp:
                    continue
                st = {"n": 0, "params": grp}
                self._hook_groups.append(st)
                st["handles"] = [p.register_post_accumulate_grad_hook(lambda _p, st=st: self._arrived(st)) for p in grp]
        self._init_step_counter()

    def _arrived(self, st) -> None:
        st["n"] += 1
        if st["n"] == len(st["params"]):
            st["n"] = 0
            self._sink(st["params"], [p.grad for p in st["params"]], None, from_hook=True)

    def _init_step_counter(self):
        """the device step counter must exist before any CUDA-graph capture (a tensor created inside a capture is
        re-initialised by every replay)."""
        p0 = self.param_groups[0]["params"][0]
        if p0.is_cuda:
            st = self.state.get(p0, {})
            s0 = int(st["step"].item()) if "step" in st else 0
            if self._step_dev is None:
                self._step_dev = torch.full((1,), s0, dtype=torch.int32, device=p0.device)
            else:                               # keep the tensor a captured graph already references
                self._step_dev.fill_(s0)
            if self._hyper_dev is None:
                ng = len(self.param_groups)
                self._hyper_host = torch.full((ng, 2), float("nan"), dtype=torch.float32).pin_memory()
                self._hyper_dev = torch.zeros(ng, 2, dtype=torch.float32, device=p0.device)
                self.refresh_hyper()

    def refresh_hyper(self) -> None:
        """upload lr / grad_scale of every group if they changed since the last upload (stream-ordered copy on the
        current stream).  Never called while capturing: a graph reads the values its caller uploaded before replay."""
        if self._hyper_dev is None:
            return
        if self._hyper_dev.shape[0] != len(self.param_groups):      # add_param_group after construction
            ng = len(self.param_groups)
            self._hyper_host = torch.full((ng, 2), float("nan"), dtype=torch.float32).pin_memory()
            self._hyper_dev = torch.zeros(ng, 2, dtype=torch.float32, device=self._hyper_dev.device)
        changed = False
        for gi, group in enumerate(self.param_groups):
            lr, gs = float(group["lr"]), float(self.grad_scale)
            if self._hyper_host[gi, 0].item() != lr or self._hyper_host[gi, 1].item() != gs:
                changed = True
        if changed:
            # a fresh pinned buffer per change: an earlier asynchronous copy may not have read the old one yet
            host = torch.empty_like(self._hyper_host).pin_memory()
            for gi, group in enumerate(self.param_groups):
                host[gi, 0] = float(group["lr"])
                host[gi, 1] = float(self.grad_scale)
            self._hyper_host = host
            self._hyper_dev.copy_(host, non_blocking=True)

    def _state_of(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32)            # torch.optim.Adam layout
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def init_state(self) -> None:
        """allocate exp_avg / exp_avg_sq of every parameter now.  Must run before a CUDA-graph capture of step():
        state created inside a capture lives in the graph's pool and would be re-zeroed by every replay."""
        for group in self.param_groups:
            for p in group["params"]:
                if p.requires_grad:
                    self._state_of(p)
        if self._step_dev is None:
            self._init_step_counter()

    def _table(self, gi, plist, grads, chunk: int = CHUNK):
        key = (weight_cache.generation, chunk, tuple((p.data_ptr(), g.data_ptr()) for p, g in zip(plist, grads)))
        ent = self._tables.get(gi)
        if ent is not None and ent[0] == key:
            return ent[1], ent[2]
        rows = []
        refreshed = set()
        for p, g in zip(plist, grads):
            st = self._state_of(p)
            bv = weight_cache.bf16_view(p)          # tensor-core operand copy, rewritten by the same kernel
            bptr = 0
            if bv is not None and bv.is_contiguous() and bv.numel() == p.numel():
                bptr = bv.data_ptr()
                refreshed.add(id(p))
            if p.dtype != torch.float32 or not p.is_contiguous() or not g.is_contiguous() or \
                    g.dtype not in (torch.float32, torch.bfloat16):
                raise RuntimeError("FusedAdam needs contiguous float32 parameters and float32 / bfloat16 gradients")
            n = p.numel()
            gsz = g.element_size()
            gate = self._row_gate.get(id(p))
            if gate is not None and self.param_groups[self._group_of[id(p)]]["weight_decay"] == 0 and p.dim() == 2 and \
                    p.shape[1] % 4 == 0:
                # row-gated chunks (whole rows): rows whose gradient has been zero in every step are skipped
                live, rl = gate["live"], p.shape[1]
                rpc = GATED_ROWS
                for r0 in range(0, p.shape[0], rpc):
                    nr = min(rpc, p.shape[0] - r0)
                    off = r0 * rl
                    rows.append((p.data_ptr() + 4 * off, st["exp_avg"].data_ptr() + 4 * off,
                                 st["exp_avg_sq"].data_ptr() + 4 * off, g.data_ptr() + gsz * off,
                                 (bptr + 2 * off) if bptr else 0, nr * rl, 1 if gsz == 2 else 0, live.data_ptr() + r0, rl))
                continue
            for off in range(0, n, chunk):
                cnt = min(chunk, n - off)
                rows.append((p.data_ptr() + 4 * off, st["exp_avg"].data_ptr() + 4 * off, st["exp_avg_sq"].data_ptr() + 4 * off,
                             g.data_ptr() + gsz * off, (bptr + 2 * off) if bptr else 0, cnt, 1 if gsz == 2 else 0, 0, 0))
        host = torch.from_numpy(np.array(rows, dtype=_DESC).view(np.uint8).reshape(-1)).pin_memory()
        dev = torch.empty(host.numel(), dtype=torch.uint8, device=plist[0].device)
        dev.copy_(host, non_blocking=True)
        self._tables[gi] = (key, dev, len(rows))
        self._refreshed[gi] = refreshed
        # a table uploaded inside a CUDA-graph capture is referenced by the graph's memcpy node for as long as the graph
        # lives; an eager rebuild (new gradient buffers) only has to outlive the launch that follows, so it replaces the
        # previous eager table of the same key instead of accumulating
        if torch.cuda.is_current_stream_capturing():
            self._keepalive.append((host, dev))
        else:
            self._eager_tables[gi] = (host, dev)
        return dev, len(rows)

    def _advance(self, device):
        if self._step_dev is None or self._hyper_dev is None:
            self._init_step_counter()
        if not self._stepped:
            if not torch.cuda.is_current_stream_capturing():
                self.refresh_hyper()
            self._step_dev += 1
            self._stepped = True

    def register_row_sparse(self, param: torch.nn.Parameter, ids_fn) -> None:
        """`param` is an embedding table [V, H] whose gradient is zero outside the rows ``ids_fn()`` returns at update time
        (int64, any shape, duplicates allowed; a fixed device tensor when the step is graph-captured; under data
        parallelism the ids of EVERY rank -- `LayerwiseReducer.gathered_ids`): BertEmbeddings.word_embeddings with the
        batch's input_ids (models/mmbert.py:52-63).  The optimizer keeps one byte per row, set for every id it has
        ever been shown; rows never shown have zero gradient and zero moments, so torch.optim.Adam would leave them
        bit-identical -- the kernel skips them (weight_decay == 0 only; with weight decay every row moves and the gate
        is ignored).  Loading a state dict marks every row live."""
        if param.dim() != 2:
            raise ValueError("register_row_sparse: a 2-D embedding table is expected")
        self._row_gate[id(param)] = {"live": torch.zeros(param.shape[0], dtype=torch.uint8, device=param.device),
                                     "ids_fn": ids_fn}
        self._tables.clear()

    def _launch(self, gi, key, plist, grads, max_ctas: int = 0, background: bool = False, fast: bool = False):
        group = self.param_groups[gi]
        for p in plist:
            gate = self._row_gate.get(id(p))
            if gate is not None:
                ops.mark_rows(gate["live"], gate["ids_fn"]())
        # bulk updates (a whole layer under the backward pass, or a plain step() over millions of parameters) stream best with
        # long one-group CTAs (0.95 of the HBM peak over 91 M parameters; the two-group loop with 8192-element chunks reaches
        # 0.74); the two-group loop is for the SMALL updates on the critical path, where per-CTA latency is the cost
        # (row-gated tables do not count: only their live rows are touched)
        # `fast`: a bulk update with nothing left to hide under (the last encoder layer of the backward pass): the gentle
        # loop has a latency floor of ~32 dependent round trips per CTA (68 us for one layer inside the step)
        gentle = (background or sum(p.numel() for p in plist if id(p) not in self._row_gate) >= SMALL_UPDATE) and not fast
        table, n = self._table(key, plist, grads, CHUNK_BACKGROUND if gentle else CHUNK)
        b1, b2 = group["betas"]
        ops.adam_step_dev(table, n, self._hyper_dev[gi], b1, b2, group["eps"], group["weight_decay"], self._step_dev,
                          max_ctas, gentle)

    @torch.no_grad()
    def _sink(self, params, grads, side_stream, tail: bool = False, from_hook: bool = False) -> bool:
        """functional.set_grad_sink target (and the early_groups hook): update one set of parameters now, on the
        optimizer stream; with reduce_fn the exchange runs on a communication stream ahead of it, so the all-reduce
        of the next set overlaps this set's update."""
        gis = {self._group_of.get(id(p)) for p in params}
        if len(gis) != 1 or None in gis or any(id(p) in self._early_ids for p in params) or \
                (not from_hook and any(p.grad is not None for p in params)):
            return False            # unknown / shared / accumulating parameters: leave them to autograd + step()
        gi = gis.pop()
        gs = [g.contiguous() for g in grads]
        for p, g in zip(params, gs):
            if not from_hook:
                p.grad = g          # visible to hooks / loggers exactly as after a normal backward
            self._early_ids.add(id(p))
        if from_hook or self._sink_schedule == [1]:
            self._flush_early(gi, list(params), gs, side_stream, background=not from_hook, fast=tail and self.reduce_fn is None)
        else:
            self._pending.append((gi, list(params), gs))
            if len(self._pending) >= self._sink_schedule[min(self._sink_round, len(self._sink_schedule) - 1)]:
                self._sink_round += 1
                self._flush_pending(side_stream)
        return True

    def _flush_pending(self, side_stream) -> None:
        if not self._pending:
            return
        gi = self._pending[0][0]
        params = [p for _, ps, _ in self._pending for p in ps]
        grads = [g for _, _, gs in self._pending for g in gs]
        self._pending = []
        self._flush_early(gi, params, grads, side_stream)

    def _flush_early(self, gi, params, gs, side_stream, background: bool = True, fast: bool = False) -> None:
        """exchange (reduce_fn, communication stream) + update (optimizer stream) of `params`, ordered after everything
        enqueued so far on the current stream and on `side_stream`."""
        dev = params[0].device
        main = torch.cuda.current_stream(dev)
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream(dev)
            self._comm_stream = torch.cuda.Stream(dev)
            self._hook_stream = torch.cuda.Stream(dev, priority=-1)   # small updates: ahead of the bulk update's CTAs
        self._advance(dev)          # on the main stream: ordered before every update of this iteration
        # hook groups (embeddings at the tail of the step, heads) update on their own stream: queued behind the long
        # background update of the last encoder layer they would end ~50 us after the gradients are there
        upd = self._opt_stream if background else self._hook_stream
        first = self._comm_stream if self.reduce_fn is not None else upd
        ev = torch.cuda.Event()
        ev.record(main)
        first.wait_event(ev)
        if side_stream is not None and side_stream is not main:
            ev2 = torch.cuda.Event()
            ev2.record(side_stream)
            first.wait_event(ev2)
        gr = gs
        if self.reduce_fn is not None:
            with torch.cuda.stream(self._comm_stream):
                gr = self.reduce_fn(list(params), gs)
            ev3 = torch.cuda.Event()
            ev3.record(self._comm_stream)
            upd.wait_event(ev3)
        with torch.cuda.stream(upd):
            # layers handed in by the backward node run underneath the rest of the backward pass (background); hook groups
            # (embeddings at the tail of the step, the heads) are small and on or near the critical path
            self._launch(gi, ("early", id(params[0]), len(params)), list(params), gr, self.early_ctas if background else 0,
                         background, fast)
        self._early_keep.append(gs)

    @torch.no_grad()
    def step(self, closure=None, grads=None):
        """`grads` (optional): list aligned with the parameters that have gradients, to read the gradients from
        other buffers (e.g. the all-reduced flat buckets of mmvqa_b200.parallel) instead of ``p.grad``."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if grads is not None and self._early_ids:
            raise RuntimeError("FusedAdam: step(grads=...) cannot be combined with overlap_backward")
        # layers still waiting for a full sink group (the backward node joined its side branch before returning, so
        # the current stream already orders their weight gradients)
        self._flush_pending(None)
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None and id(p) not in self._early_ids]
            if not plist:
                continue
            g = [p.grad for p in plist] if grads is None else grads
            if self.reduce_fn is not None and grads is None:
                g = self.reduce_fn(plist, g)
            self._advance(plist[0].device)
            self._launch(gi, gi, plist, g)
        if self._early_ids:         # join the optimizer stream; its gradient buffers may be released after this point
            dev = self._opt_stream.device
            for st_ in (self._opt_stream, self._hook_stream):
                ev = torch.cuda.Event()
                ev.record(st_)
                torch.cuda.current_stream(dev).wait_event(ev)
            self._early_ids = set()
            self._early_keep = []
        for st in self._hook_groups:
            st["n"] = 0
        self._sink_round = 0
        self._stepped = False
        ids = set()
        for r in self._refreshed.values():
            ids |= r
        weight_cache.note_optimizer_step(ids)
        return loss

    def close(self) -> None:
        """remove this optimizer's gradient sink (overlap_backward)."""
        if Fn.grad_sink() == self._sink:
            Fn.set_grad_sink(None)
        for st in self._hook_groups:
            for h in st["handles"]:
                h.remove()
        self._hook_groups = []

    def covers_weight_cache(self) -> bool:
        """True if every cached tensor-core operand copy is rewritten by this optimizer's kernel (then a captured
        graph of the step needs no weight casts)."""
        ids = set()
        for r in self._refreshed.values():
            ids |= r
        return bool(self._tables) and weight_cache.covered_by(ids)

    def state_dict(self):
        if self._step_dev is not None:                  # publish the device step counter in torch's layout
            s = float(self._step_dev.item())
            for st in self.state.values():
                if "step" in st:
                    st["step"] = torch.tensor(s)
        return super().state_dict()

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        for gate in self._row_gate.values():
            gate["live"].fill_(1)      # loaded moments may be non-zero anywhere
        self._tables.clear()
        self._init_step_counter()      # in place: a captured graph keeps reading the same device counter
        self.refresh_hyper()
