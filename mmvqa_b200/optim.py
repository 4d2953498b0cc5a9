"""Fused multi-tensor Adam: torch.optim.Adam semantics (vqamed2019/train.py:160) in ONE
kernel launch over every parameter, with a state_dict compatible with torch.optim.Adam so the
reference's ``recorder_2.pt`` checkpoints round-trip (pretrain/roco_train.py:164-171)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .functional import weight_cache

CHUNK = 32768
_DESC = np.dtype([("p", "<u8"), ("m", "<u8"), ("v", "<u8"), ("g", "<u8"), ("bf16_out", "<u8"), ("n", "<i8"), ("flags", "<i8")])


class FusedAdam(torch.optim.Optimizer):
    """Drop-in for ``torch.optim.Adam(params, lr)`` (no amsgrad / maximize / capturable flags).

    ``grad_scale`` multiplies every gradient inside the kernel (1/world_size after an all-reduce SUM).
    The step counter lives on the device, so a captured CUDA graph of ``step()`` can be replayed."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = 1.0
        self._tables = {}
        self._keepalive = []       # pinned host tables referenced by memcpy nodes of captured graphs
        self._refreshed = {}       # group index -> ids of parameters whose bf16 cache copy the kernel rewrites
        self._step_dev = None

    def _state_of(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32)            # torch.optim.Adam layout
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _table(self, gi, group, grads):
        plist = [p for p in group["params"] if p.grad is not None]
        key = (weight_cache.generation, tuple((p.data_ptr(), g.data_ptr()) for p, g in zip(plist, grads)))
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
            for off in range(0, n, CHUNK):
                cnt = min(CHUNK, n - off)
                rows.append((p.data_ptr() + 4 * off, st["exp_avg"].data_ptr() + 4 * off, st["exp_avg_sq"].data_ptr() + 4 * off,
                             g.data_ptr() + gsz * off, (bptr + 2 * off) if bptr else 0, cnt, 1 if gsz == 2 else 0))
        host = torch.from_numpy(np.array(rows, dtype=_DESC).view(np.uint8).reshape(-1)).pin_memory()
        dev = torch.empty(host.numel(), dtype=torch.uint8, device=plist[0].device)
        dev.copy_(host, non_blocking=True)
        self._tables[gi] = (key, dev, len(rows))
        self._refreshed[gi] = refreshed
        self._keepalive.append((host, dev))
        return dev, len(rows)

    @torch.no_grad()
    def step(self, closure=None, grads=None):
        """`grads` (optional): list aligned with the parameters that have gradients, to read the gradients from
        other buffers (e.g. the all-reduced flat buckets of mmvqa_b200.parallel) instead of ``p.grad``."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            g = [p.grad for p in plist] if grads is None else grads
            table, n = self._table(gi, group, g)
            if self._step_dev is None:
                st0 = self._state_of(plist[0])
                self._step_dev = torch.full((1,), int(st0["step"].item()), dtype=torch.int32, device=plist[0].device)
            self._step_dev += 1
            b1, b2 = group["betas"]
            ops.adam_step(table, n, group["lr"], b1, b2, group["eps"], group["weight_decay"], 0, self._step_dev,
                          self.grad_scale)
        ids = set()
        for r in self._refreshed.values():
            ids |= r
        weight_cache.note_optimizer_step(ids)
        return loss

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
        self._tables.clear()
        self._step_dev = None
