"""SERF activation, x * erf(softplus(x)).  Mirrors models/serf.py:8-24 of the reference."""
import torch
import torch.nn as nn

from .. import functional as Fn
from .._lib import ACT_SERF


class SERF(nn.Module):
    """Drop-in for the reference ``SERF(thresh=50)``: one fused kernel instead of five ATen passes
    (clamp, exp, log1p, erf, mul); backward is the closed-form derivative in one pass."""

    def __init__(self, thresh=50):
        super().__init__()
        if thresh != 50:
            raise NotImplementedError("the CUDA kernel hard-codes the reference's default clamp thresh=50")
        self.thresh = thresh

    def forward(self, x):
        return self.serf_log1pexp(x)

    def serf(self, x):
        # the reference's "naive" variant differs only for x > 50 where exp overflows; same kernel
        return self.serf_log1pexp(x)

    def serf_log1pexp(self, x):
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        return Fn.ActFn.apply(x, ACT_SERF)
