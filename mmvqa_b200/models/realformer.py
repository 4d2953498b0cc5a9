"""RealFormer (residual attention) encoder block.  Mirrors models/realformer.py of the reference."""
import torch
import torch.nn as nn

from .. import functional as Fn
from .serf import SERF


def _seed():
    return int(torch.randint(0, 2 ** 31 - 2 ** 16, (1,)).item())


def block_params(b):
    """The 10 parameters of one block in the order RealFormerEncoderFn expects."""
    return (b.kqv.weight, b.proj.weight, b.ln1.weight, b.ln1.bias, b.ff[0].weight, b.ff[0].bias, b.ff[2].weight,
            b.ff[2].bias, b.ln2.weight, b.ln2.bias)


def run_blocks(blocks, x, prev, mask, training):
    """x [B,T,H], prev in the reference layout [B,T,T,h] (or None), mask [B,T] -> (x, prev)."""
    first = blocks[0]
    for b in blocks:
        if (b.head_cnt, b.emb_s, b.dp.p, b.ff[3].p) != (first.head_cnt, first.emb_s, first.dp.p, first.ff[3].p):
            raise ValueError("fused RealFormer encoder needs identical block hyper-parameters")
    x = Fn.to_compute(x)
    maskf = None if mask is None else mask.to(torch.float32).contiguous()
    p1 = first.dp.p if training else 0.0
    p2 = first.ff[3].p if training else 0.0
    native_prev = None if prev is None else prev.permute(0, 3, 1, 2)     # kernel layout [B,h,T,T]
    params = []
    for b in blocks:
        params.extend(block_params(b))
    y, scores = Fn.RealFormerEncoderFn.apply(x, maskf, native_prev, first.head_cnt, p1, p2,
                                             _seed() if (p1 > 0 or p2 > 0) else 0, *params)
    return y, scores.permute(0, 2, 3, 1)                                   # reference layout [B,Tq,Tk,h]


class ResEncoderBlock(nn.Module):
    """realformer.py:9-51.  One [3*emb_s, emb_s] kqv weight shared by every head (split order k, q, v), no
    biases on kqv / proj, query-side mask carried in the returned pre-softmax scores, post-LN (eps 1e-5),
    SERF feed-forward.  ``forward`` returns (x, prev) with prev in the reference layout [B,T,T,h]
    (a permuted view of the kernel's [B,h,T,T] fp32 buffer)."""

    def __init__(self, emb_s=32, head_cnt=8, dp1=0.1, dp2=0.1):
        super().__init__()
        emb = emb_s * head_cnt
        self.kqv = nn.Linear(emb_s, 3 * emb_s, bias=False)
        self.dp = nn.Dropout(dp1)
        self.proj = nn.Linear(emb, emb, bias=False)
        self.head_cnt = head_cnt
        self.emb_s = emb_s
        self.ln1 = nn.LayerNorm(emb)
        self.ln2 = nn.LayerNorm(emb)
        self.ff = nn.Sequential(
            nn.Linear(emb, 4 * emb),
            SERF(),
            nn.Linear(4 * emb, emb),
            nn.Dropout(dp2),
        )

    def resmha(self, x, prev=None, mask=None):
        """Attention sub-block only: (dropout(proj(attention(x))), scores), realformer.py:30-45."""
        B, T, _ = x.shape
        x = Fn.to_compute(x)
        maskf = None if mask is None else mask.to(torch.float32).contiguous()
        native_prev = None if prev is None else prev.permute(0, 3, 1, 2)
        res, scores = Fn.RFAttentionFn.apply(x, maskf, native_prev, self.kqv.weight, self.head_cnt)
        p = self.dp.p if self.training else 0.0
        out = Fn.linear(res, self.proj.weight, None)
        if p > 0:
            out = Fn.DropoutFn.apply(out, p, _seed())
        return out, scores.permute(0, 2, 3, 1)

    def forward(self, x, prev=None, mask=None):
        return run_blocks([self], x, prev, mask, self.training)
