"""Transformer encoder layer.  Mirrors models/transformer.py of the reference (gelu,
MultiHeadedSelfAttention, PositionWiseFeedForward, BertLayer with its 4 sharing modes x pre/post norm)."""
import numpy as np
import torch
import torch.nn as nn

from .. import functional as Fn
from .._lib import ACT_GELU


def _seed():
    return int(torch.randint(0, 2 ** 31 - 1, (1,)).item())


def _maskf(mask):
    return None if mask is None else mask.to(torch.float32).contiguous()


def gelu(x):
    """exact erf GELU (transformer.py:7-8) as one kernel."""
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return Fn.ActFn.apply(x, ACT_GELU)


class MultiHeadedSelfAttention(nn.Module):
    """q/k/v projections run as ONE [M,H]x[H,3H] GEMM; QK^T, key mask, softmax, dropout and PV are one
    fused kernel per (batch, head).  ``self.scores`` holds the attention probabilities."""

    def __init__(self, args):
        super(MultiHeadedSelfAttention, self).__init__()
        self.proj_q = nn.Linear(args.hidden_size, args.hidden_size)
        self.proj_k = nn.Linear(args.hidden_size, args.hidden_size)
        self.proj_v = nn.Linear(args.hidden_size, args.hidden_size)
        self.drop = nn.Dropout(args.hidden_dropout_prob)
        self.scores = None
        self.n_heads = args.heads

    def forward(self, x, mask):
        x = Fn.to_compute(x)
        p = self.drop.p if self.training else 0.0
        h, probs = Fn.MHSAFn.apply(x, _maskf(mask), self.proj_q.weight, self.proj_q.bias, self.proj_k.weight,
                                   self.proj_k.bias, self.proj_v.weight, self.proj_v.bias, self.n_heads, p,
                                   _seed() if p > 0 else 0)
        self.scores = probs
        return h

    def split_last(self, x, shape):
        shape = list(shape)
        assert shape.count(-1) <= 1
        if -1 in shape:
            shape[shape.index(-1)] = int(x.size(-1) / -np.prod(shape))
        return x.view(*x.size()[:-1], *shape)

    def merge_last(self, x, n_dims):
        s = x.size()
        assert n_dims > 1 and n_dims < len(s)
        return x.view(*s[:-n_dims], -1)


class PositionWiseFeedForward(nn.Module):
    """fc2(gelu(fc1(x))): bias + erf-GELU live in the fc1 GEMM epilogue."""

    def __init__(self, args):
        super(PositionWiseFeedForward, self).__init__()
        self.fc1 = nn.Linear(args.hidden_size, args.hidden_size * 4)
        self.fc2 = nn.Linear(args.hidden_size * 4, args.hidden_size)

    def forward(self, x, residual=None, dropout_p=0.0):
        x = Fn.to_compute(x)
        h = Fn.linear(x, self.fc1.weight, self.fc1.bias, act=ACT_GELU)
        return Fn.linear(h, self.fc2.weight, self.fc2.bias, residual=residual, dropout_p=dropout_p,
                         seed=_seed() if dropout_p > 0 else 0)


class BertLayer(nn.Module):
    """transformer.py:50-97.  norm1 is shared by every layer and, in pre-norm mode, used for BOTH
    sub-blocks (norm2 stays unused) -- that is the reference's behaviour and is kept.  The residual add
    and dropout of each sub-block are fused into the epilogue of its last GEMM."""

    def __init__(self, args, share='all', norm='pre'):
        super(BertLayer, self).__init__()
        self.share = share
        self.norm_pos = norm
        self.norm1 = nn.LayerNorm(args.hidden_size, eps=1e-12)
        self.norm2 = nn.LayerNorm(args.hidden_size, eps=1e-12)
        self.drop1 = nn.Dropout(args.hidden_dropout_prob)
        self.drop2 = nn.Dropout(args.hidden_dropout_prob)
        if self.share == 'ffn':
            self.attention = nn.ModuleList([MultiHeadedSelfAttention(args) for _ in range(args.n_layers)])
            self.proj = nn.ModuleList([nn.Linear(args.hidden_size, args.hidden_size) for _ in range(args.n_layers)])
            self.feedforward = PositionWiseFeedForward(args)
        elif self.share == 'att':
            self.attention = MultiHeadedSelfAttention(args)
            self.proj = nn.Linear(args.hidden_size, args.hidden_size)
            self.feedforward = nn.ModuleList([PositionWiseFeedForward(args) for _ in range(args.n_layers)])
        elif self.share == 'all':
            self.attention = MultiHeadedSelfAttention(args)
            self.proj = nn.Linear(args.hidden_size, args.hidden_size)
            self.feedforward = PositionWiseFeedForward(args)
        elif self.share == 'none':
            self.attention = nn.ModuleList([MultiHeadedSelfAttention(args) for _ in range(args.n_layers)])
            self.proj = nn.ModuleList([nn.Linear(args.hidden_size, args.hidden_size) for _ in range(args.n_layers)])
            self.feedforward = nn.ModuleList([PositionWiseFeedForward(args) for _ in range(args.n_layers)])

    def _pick(self, mod, layer_num):
        return mod[layer_num] if isinstance(mod, nn.ModuleList) else mod

    def forward(self, hidden_states, attention_mask, layer_num):
        x = Fn.to_compute(hidden_states)
        att = self._pick(self.attention, layer_num)
        proj = self._pick(self.proj, layer_num)
        ffn = self._pick(self.feedforward, layer_num)
        p1 = self.drop1.p if self.training else 0.0
        p2 = self.drop2.p if self.training else 0.0
        n1, n2 = self.norm1, self.norm2
        out = x
        if self.norm_pos == 'pre':
            a = att(Fn.add_layer_norm(x, None, n1.weight, n1.bias, n1.eps), attention_mask)
            out = Fn.linear(a, proj.weight, proj.bias, residual=x, dropout_p=p1, seed=_seed() if p1 > 0 else 0)
            out = ffn(Fn.add_layer_norm(out, None, n1.weight, n1.bias, n1.eps), residual=out, dropout_p=p2)
        if self.norm_pos == 'post':
            a = att(x, attention_mask)
            s = Fn.linear(a, proj.weight, proj.bias, residual=x, dropout_p=p1, seed=_seed() if p1 > 0 else 0)
            out = Fn.add_layer_norm(s, None, n1.weight, n1.bias, n1.eps)
            s = ffn(out, residual=out, dropout_p=p2)
            out = Fn.add_layer_norm(s, None, n2.weight, n2.bias, n2.eps)
        return out
