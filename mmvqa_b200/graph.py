"""CUDA-graph capture of a whole training step (forward + loss + backward + optimizer).

At the flagship VQA shape (B=16, T=28) a step is ~300 short kernels; launching them from Python
costs more than running them, so the step is captured once and replayed (no tracing compiler: the
graph records exactly the C-ABI launches the eager path makes)."""
from __future__ import annotations

from typing import Callable, Sequence

import torch

from . import _lib
from .functional import invalidate_weight_cache


class GraphedTrainStep:
    """``loss_fn(*inputs) -> scalar loss`` is run as zero_grad -> forward -> backward -> [post_backward] ->
    [eager_between] -> optimizer.step().

    Without ``eager_between`` the whole step is ONE graph.  With it (data parallel: the NCCL gradient all-reduce)
    the step is two graphs -- forward/backward/bucket-packing and the optimizer -- with the collective launched
    eagerly between the two replays, so no NCCL kernel is ever recorded into a graph.
    ``replay(*inputs)`` copies new inputs into the static buffers (device or pinned-host sources, non-blocking),
    replays and returns the static loss tensor.

    What a replay must NOT freeze: (1) dropout masks -- the graph increments a device step counter that every dropout
    kernel mixes into its seed (``mmvqa_set_dropout_counter``), so each replay draws new masks and forward / backward of
    one replay agree; (2) ``lr`` / ``grad_scale`` -- ``FusedAdam`` reads them from device memory that
    ``replay`` refreshes from ``param_groups`` first, so ``ReduceLROnPlateau.step()`` between replays takes effect
    (vqamed2019/train.py:160-161).  ``close()`` drops the graphs (needed before an NCCL communicator they captured is
    destroyed)."""

    def __init__(self, loss_fn: Callable, example_inputs: Sequence[torch.Tensor], optimizer, warmup: int = 3,
                 post_backward: Callable = None, eager_between: Callable = None, step_kwargs: Callable = None,
                 capture_error_mode: str = "global", main_priority: int = 0):
        self.loss_fn = loss_fn
        self.optimizer = optimizer
        self.post_backward = post_backward      # capturable, e.g. GradBuckets.pack
        self.eager_between = eager_between      # not captured, e.g. GradBuckets.allreduce
        self.step_kwargs = step_kwargs          # e.g. lambda: dict(grads=buckets.grads(plist))
        # the static inputs are views into ONE flat buffer (256-byte aligned slots): a caller that keeps its batches in
        # the same packed layout (`pack_like`) refreshes all of them with a single copy (`replay_packed`) instead of
        # one copy per tensor -- nine launches and ~35 us per step at the flagship shape
        offs, total = [], 0
        for t in example_inputs:
            offs.append(total)
            total += (t.numel() * t.element_size() + 255) // 256 * 256
        self._offsets, self._flat_bytes = offs, total
        self.static_flat = torch.empty(total, dtype=torch.uint8, device=example_inputs[0].device)
        self.static_inputs = self._views(self.static_flat, example_inputs)
        for dst, src in zip(self.static_inputs, example_inputs):
            dst.copy_(src)
        if hasattr(optimizer, "init_state"):
            optimizer.init_state()              # optimizer state must not be born inside the capture
        if hasattr(optimizer, "refresh_hyper"):
            optimizer.refresh_hyper()
        # device step counter of the dropout masks: registered for the warm-up and the capture only, so the pointer is
        # baked into this graph's kernels and eager calls made later keep drawing host seeds
        self._drop_ctr = torch.zeros(1, dtype=torch.int64, device=example_inputs[0].device)
        _lib.set_dropout_counter(self._drop_ctr)
        try:
            self._capture(warmup, capture_error_mode, main_priority)
        finally:
            _lib.set_dropout_counter(None)

    def _capture(self, warmup, capture_error_mode, main_priority):
        optimizer = self.optimizer
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                if self.eager_between is not None:
                    self.eager_between()
                self._opt()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        covers = getattr(self.optimizer, "covers_weight_cache", None)
        self.casts_in_graph = not (covers is not None and covers())
        if self.casts_in_graph:
            invalidate_weight_cache()      # weight casts must be recorded inside the graph
        self.optimizer.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        self.graph_opt = None
        n0 = _lib.launch_count()
        # main_priority < 0: the step is captured on a high-priority stream, so the kernels of the main chain win the
        # SM slots over the work forked onto default-priority streams (weight-gradient branch, optimizer stream)
        cap_stream = torch.cuda.Stream(priority=main_priority) if main_priority != 0 else None
        from . import functional as _Fn
        _Fn.TRAIN_STEP_CAPTURE = True
        try:
            self._capture_graphs(cap_stream, capture_error_mode)
        finally:
            _Fn.TRAIN_STEP_CAPTURE = False
        self.launches_per_step = _lib.launch_count() - n0
        if self.casts_in_graph:
            invalidate_weight_cache()

    def _capture_graphs(self, cap_stream, capture_error_mode):
        if self.eager_between is None:
            # capture_error_mode="thread_local": needed when collectives are captured (the NCCL watchdog thread
            # polls events of earlier eager collectives while this thread captures)
            with torch.cuda.graph(self.graph, stream=cap_stream, capture_error_mode=capture_error_mode):
                self.static_loss = self._fwd_bwd(zero=False)
                self._opt()
        else:
            with torch.cuda.graph(self.graph):
                self.static_loss = self._fwd_bwd(zero=False)
            self.eager_between()
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt, pool=self.graph.pool()):
                self._opt()

    def _fwd_bwd(self, zero: bool = True):
        if zero:
            self.optimizer.zero_grad(set_to_none=True)
        self._drop_ctr.add_(1)                  # captured: every replay draws new dropout masks
        loss = self.loss_fn(*self.static_inputs)
        loss.backward()
        if self.post_backward is not None:
            self.post_backward()
        return loss.detach()

    def _opt(self):
        self.optimizer.step(**(self.step_kwargs() if self.step_kwargs is not None else {}))

    def _views(self, flat, like):
        return [flat[o:o + t.numel() * t.element_size()].view(t.dtype).view(t.shape) for o, t in zip(self._offsets, like)]

    def pack_like(self, tensors, device=None, pin_memory=False):
        """copy `tensors` (one batch, same shapes / dtypes as the example inputs) into a new flat buffer with the static
        layout; returns (flat uint8 buffer, views).  device=None keeps the batch on the host (optionally pinned)."""
        dev = device if device is not None else torch.device("cpu")
        flat = torch.empty(self._flat_bytes, dtype=torch.uint8, device=dev)
        if pin_memory and dev.type == "cpu":
            flat = flat.pin_memory()
        views = self._views(flat, tensors)
        for dst, src in zip(views, tensors):
            dst.copy_(src)
        return flat, views

    def replay_packed(self, flat):
        """replay with a batch stored in the packed layout (device or pinned host): ONE copy refreshes every input."""
        if flat is not self.static_flat:
            self.static_flat.copy_(flat, non_blocking=True)
        self._pre_replay()
        self.graph.replay()
        if self.graph_opt is not None:
            self.eager_between()
            self.graph_opt.replay()
        return self.static_loss

    def replay(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self._pre_replay()
        self.graph.replay()
        if self.graph_opt is not None:
            self.eager_between()
            self.graph_opt.replay()
        return self.static_loss

    def _pre_replay(self):
        if self.graph is None:
            raise RuntimeError("GraphedTrainStep.replay() after close()")
        if hasattr(self.optimizer, "refresh_hyper"):
            self.optimizer.refresh_hyper()      # lr / grad_scale changed by a scheduler since the last replay

    def close(self) -> None:
        """drop the captured graphs (and the NCCL work they hold) after the device has drained."""
        torch.cuda.synchronize()
        self.graph = None
        self.graph_opt = None
