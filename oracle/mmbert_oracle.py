"""CPU oracle for the MMBERT fusion-encoder hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement, in plain torch ops on CPU tensors, of the
reference algorithm (DannielSilva/MM-VQA, ``models/*.py``).  It is *not* part of the
product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or
as the timed CPU baseline.  Nothing under ``mmvqa_b200/`` imports it.

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against the reference *itself*: ``oracle/gen_golden.py`` imports the
unmodified reference modules from /root/reference in the build container, runs them on
seeded inputs and commits the input/output vectors under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against those vectors.

Everything is written as stateless functions over a flat ``params`` dict that uses the
reference's own state-dict key names, so the same dict can be loaded
into the reference modules, into this oracle and into the CUDA-backed modules.

All functions are dtype generic (fp32 for parity with the reference, fp64 for
known-answer tests).  Dropout is always off (SURVEY.md section 7 hard part 6).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

SERF_THRESH = 50.0


# ----------------------------------------------------------------------------------
# activations
# ----------------------------------------------------------------------------------
def serf(x: Tensor, thresh: float = SERF_THRESH) -> Tensor:
    """x * erf(softplus(min(x, thresh))).  Reference: models/serf.py:23-24."""
    sp = torch.log1p(torch.exp(torch.clamp(x, max=thresh)))
    return x * torch.erf(sp)


def gelu_erf(x: Tensor) -> Tensor:
    """Exact (erf) GELU.  Reference: models/transformer.py:7-8."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def activation(name: str, x: Tensor) -> Tensor:
    if name == "serf":
        return serf(x)
    if name == "gelu":
        return gelu_erf(x)
    if name == "relu":
        return torch.relu(x)
    if name == "none":
        return x
    raise ValueError(name)


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


# ----------------------------------------------------------------------------------
# a1: visual-token projector   (models/image_encoding.py:43-62, 71-87, 100-115)
# ----------------------------------------------------------------------------------
VISTOK_CONV_NAMES = ("conv2", "conv3", "conv4", "conv5", "conv7")


def vistok_project(feats: Sequence[Tensor], conv_w: Sequence[Tensor], act: str = "serf") -> List[Tensor]:
    """For every level l: v_l[b, :] = mean_{h,w} act(W_l . f_l[b, :, h, w]).

    ``feats[l]`` is NCHW, ``conv_w[l]`` is the 1x1 conv weight [hidden, C_l, 1, 1] (no bias).
    Activation is applied BEFORE pooling (image_encoding.py:74 / :103).
    """
    out = []
    for f, w in zip(feats, conv_w):
        B, C, Hh, Ww = f.shape
        pix = f.permute(0, 2, 3, 1).reshape(B, Hh * Ww, C)           # [B, HW, C]
        y = pix @ w.reshape(w.shape[0], C).t()                        # [B, HW, hidden]
        out.append(activation(act, y).mean(dim=1))
    return out


# ----------------------------------------------------------------------------------
# a2: BERT embeddings + visual-token scatter   (models/mmbert.py:52-67, HF BertEmbeddings)
# ----------------------------------------------------------------------------------
def bert_embeddings(input_ids: Tensor, token_type_ids: Tensor, word: Tensor, pos: Tensor, typ: Tensor,
                    ln_w: Tensor, ln_b: Tensor, eps: float = 1e-12) -> Tensor:
    """word[ids] + type[seg] + pos[0..T) then LayerNorm(eps=1e-12); dropout off."""
    T = input_ids.shape[1]
    e = word[input_ids] + typ[token_type_ids] + pos[:T].unsqueeze(0)
    return layer_norm(e, ln_w, ln_b, eps)


def fuse_visual_tokens(h: Tensor, vis: Sequence[Tensor]) -> Tensor:
    """h[b, n, :] = vis[n][b] for n < len(vis): overwrites positions 0..num_vis-1 (incl. [CLS]).

    Reference: models/mmbert.py:64-66 (Python double loop).  Done out of place here so
    autograd sees both the embedding path and the visual-token path.
    """
    nv = len(vis)
    if nv == 0:
        return h
    v = torch.stack(list(vis), dim=1)                                 # [B, nv, H]
    return torch.cat([v, h[:, nv:, :]], dim=1)


def prepare_input(feats, input_ids, token_type_ids, p: Params, act: str, prefix: str = "transformer.") -> Tensor:
    convs = [p[f"{prefix}trans.{n}.weight"] for n in VISTOK_CONV_NAMES][: len(feats)]
    vis = vistok_project(feats, convs, act)
    e = prefix + "bert_embedding."
    h = bert_embeddings(input_ids, token_type_ids, p[e + "word_embeddings.weight"],
                        p[e + "position_embeddings.weight"], p[e + "token_type_embeddings.weight"],
                        p[e + "LayerNorm.weight"], p[e + "LayerNorm.bias"])
    return fuse_visual_tokens(h, vis)


# ----------------------------------------------------------------------------------
# a3-a6: Transformer encoder   (models/transformer.py)
# ----------------------------------------------------------------------------------
def mhsa(x: Tensor, mask: Optional[Tensor], wq, bq, wk, bk, wv, bv, heads: int) -> Tuple[Tensor, Tensor]:
    """Multi-head self-attention with a KEY-side additive mask.  transformer.py:19-30.

    Returns (merged head outputs [B,T,H], probabilities [B,heads,T,T])."""
    B, T, H = x.shape
    d = H // heads
    q = (x @ wq.t() + bq).view(B, T, heads, d).transpose(1, 2)
    k = (x @ wk.t() + bk).view(B, T, heads, d).transpose(1, 2)
    v = (x @ wv.t() + bv).view(B, T, heads, d).transpose(1, 2)
    s = q @ k.transpose(-2, -1) / math.sqrt(d)
    if mask is not None:
        s = s - 10000.0 * (1.0 - mask[:, None, None, :].to(s.dtype))
    pr = torch.softmax(s, dim=-1)
    o = (pr @ v).transpose(1, 2).reshape(B, T, H)
    return o, pr


def ffn_gelu(x, w1, b1, w2, b2):
    """fc2(gelu(fc1(x))).  transformer.py:42-48."""
    return gelu_erf(x @ w1.t() + b1) @ w2.t() + b2


def bert_layer(x: Tensor, mask, p: Params, layer: int, heads: int, share: str = "none", norm: str = "pre",
               prefix: str = "transformer.blocks.") -> Tensor:
    """One BertLayer step (transformer.py:75-97).  norm1 is shared by all layers and, in
    pre-norm mode, used for BOTH sub-blocks (norm2 unused) -- that is the spec."""
    att_shared = share in ("att", "all")
    ffn_shared = share in ("ffn", "all")
    a = prefix + ("attention." if att_shared else f"attention.{layer}.")
    pj = prefix + ("proj." if att_shared else f"proj.{layer}.")
    f = prefix + ("feedforward." if ffn_shared else f"feedforward.{layer}.")
    n1w, n1b = p[prefix + "norm1.weight"], p[prefix + "norm1.bias"]
    n2w, n2b = p[prefix + "norm2.weight"], p[prefix + "norm2.bias"]

    def attn(z):
        o, _ = mhsa(z, mask, p[a + "proj_q.weight"], p[a + "proj_q.bias"], p[a + "proj_k.weight"],
                    p[a + "proj_k.bias"], p[a + "proj_v.weight"], p[a + "proj_v.bias"], heads)
        return o @ p[pj + "weight"].t() + p[pj + "bias"]

    def ff(z):
        return ffn_gelu(z, p[f + "fc1.weight"], p[f + "fc1.bias"], p[f + "fc2.weight"], p[f + "fc2.bias"])

    if norm == "pre":
        x = x + attn(layer_norm(x, n1w, n1b, 1e-12))
        x = x + ff(layer_norm(x, n1w, n1b, 1e-12))
        return x
    x = layer_norm(x + attn(x), n1w, n1b, 1e-12)
    x = layer_norm(x + ff(x), n2w, n2b, 1e-12)
    return x


def transformer_encoder(h: Tensor, mask, p: Params, n_layers: int, heads: int,
                        prefix: str = "transformer.blocks.") -> Tensor:
    """mmbert.py:90-94: share='none', norm='pre', no final LayerNorm."""
    for i in range(n_layers):
        h = bert_layer(h, mask, p, i, heads, "none", "pre", prefix)
    return h


# ----------------------------------------------------------------------------------
# a7-a9: RealFormer encoder   (models/realformer.py, mmbert.py:96-108)
# ----------------------------------------------------------------------------------
def realformer_attention(x: Tensor, prev: Optional[Tensor], mask: Optional[Tensor], kqv_w: Tensor,
                         heads: int) -> Tuple[Tensor, Tensor]:
    """Residual attention (realformer.py:30-44), without the out-projection.

    * one [3d, d] weight shared by every head, split order k, q, v (:33)
    * S = q k^T / sqrt(d) (+ prev), layout [B, Tq, Tk, heads] (:35-37)
    * mask subtracts 10000 along the QUERY axis i (:38-41) and is carried in S
    * softmax over keys, dim=2 (:43); padded keys are attended
    Returns (head-major merged output [B,T,H], S [B,T,T,heads]).
    """
    B, T, H = x.shape
    d = H // heads
    xh = x.reshape(B, T, heads, d)
    kqv = xh @ kqv_w.t()                                              # [B,T,h,3d]
    k, q, v = kqv[..., :d], kqv[..., d:2 * d], kqv[..., 2 * d:]
    s = torch.einsum("bihk,bjhk->bijh", q, k) / math.sqrt(d)
    if prev is not None:
        s = s + prev
    if mask is not None:
        s = s - 10000.0 * (1.0 - mask.to(s.dtype))[:, :, None, None]
    att = torch.softmax(s, dim=2)
    o = torch.einsum("btih,bihs->bths", att, v).reshape(B, T, H)
    return o, s


def realformer_block(x: Tensor, prev, mask, p: Params, prefix: str, heads: int = 8) -> Tuple[Tensor, Tensor]:
    """ResEncoderBlock.forward (realformer.py:47-51): post-LN, eps 1e-5, FF with SERF."""
    o, s = realformer_attention(x, prev, mask, p[prefix + "kqv.weight"], heads)
    x = layer_norm(x + o @ p[prefix + "proj.weight"].t(), p[prefix + "ln1.weight"], p[prefix + "ln1.bias"], 1e-5)
    f = serf(x @ p[prefix + "ff.0.weight"].t() + p[prefix + "ff.0.bias"]) @ p[prefix + "ff.2.weight"].t() \
        + p[prefix + "ff.2.bias"]
    x = layer_norm(x + f, p[prefix + "ln2.weight"], p[prefix + "ln2.bias"], 1e-5)
    return x, s


def realformer_encoder(h: Tensor, mask, p: Params, n_layers: int, heads: int = 8,
                       prefix: str = "transformer.mains.") -> Tensor:
    """mmbert.py:103-108: thread prev through the layers, drop the last one."""
    prev = None
    for i in range(n_layers):
        h, prev = realformer_block(h, prev, mask, p, f"{prefix}{i}.", heads)
    return h


# ----------------------------------------------------------------------------------
# a11: heads   (models/mmbert.py:129-172)
# ----------------------------------------------------------------------------------
def mean_pooling(h: Tensor, mask: Tensor) -> Tensor:
    """Mask-weighted mean over tokens, denominator clamped at 1e-9.  mmbert.py:169-172."""
    m = mask.unsqueeze(-1).to(h.dtype)
    return (h * m).sum(1) / torch.clamp(m.sum(1).expand(-1, h.shape[-1]), min=1e-9)


def classifier_head(z: Tensor, p: Params) -> Tensor:
    """classifier(SERF(fc1(z))): Linear -> LN(1e-12) -> Linear, no act in between. mmbert.py:133-137."""
    z = serf(z @ p["fc1.weight"].t() + p["fc1.bias"])
    z = z @ p["classifier.0.weight"].t() + p["classifier.0.bias"]
    z = layer_norm(z, p["classifier.1.weight"], p["classifier.1.bias"], 1e-12)
    return z @ p["classifier.2.weight"].t() + p["classifier.2.bias"]


def supcon_head(z: Tensor, p: Params) -> Tensor:
    """normalize(head(z)): Linear -> SERF -> Linear(H,128), L2 normalise. mmbert.py:143-148,157."""
    z = serf(z @ p["head.0.weight"].t() + p["head.0.bias"])
    z = z @ p["head.2.weight"].t() + p["head.2.bias"]
    return z / torch.clamp(z.norm(dim=1, keepdim=True), min=1e-12)


def encode(feats, input_ids, segment_ids, input_mask, p: Params, *, encoder: str, n_layers: int, heads: int,
           act: str = "serf") -> Tensor:
    h = prepare_input(feats, input_ids, segment_ids, p, act)
    if encoder == "realformer":
        return realformer_encoder(h, input_mask, p, n_layers)
    return transformer_encoder(h, input_mask, p, n_layers, heads)


def model_forward(feats, input_ids, segment_ids, input_mask, p: Params, *, encoder: str, n_layers: int,
                  heads: int = 12, act: str = "serf", dataset: str = "VQA-Med", supcon: bool = False):
    """Model.forward from feature maps on (mmbert.py:150-167).

    VQA-Med -> logits [B, C];  roco/MLM -> logits [B,T,V] (and feat [B,128] if supcon)."""
    h = encode(feats, input_ids, segment_ids, input_mask, p, encoder=encoder, n_layers=n_layers, heads=heads,
               act=act)
    if dataset == "VQA-Med":
        return classifier_head(mean_pooling(h, input_mask), p)
    logits = classifier_head(h, p)
    if supcon:
        return logits, supcon_head(mean_pooling(h, input_mask), p)
    return logits


# ----------------------------------------------------------------------------------
# a12-a14: losses
# ----------------------------------------------------------------------------------
def asl_single_label(logits: Tensor, target: Tensor, gamma_pos: float = 0.0, gamma_neg: float = 4.0,
                     eps: float = 0.1, reduction: str = "mean") -> Tensor:
    """ASLSingleLabel.forward (models/asl_singlelabel.py:23-52)."""
    C = logits.shape[-1]
    lp = torch.log_softmax(logits, dim=-1)
    t = torch.zeros_like(logits).scatter_(1, target.long().unsqueeze(1), 1.0)
    pr = torch.exp(lp)
    base = 1.0 - pr * t - (1.0 - pr) * (1.0 - t)
    w = torch.pow(base, gamma_pos * t + gamma_neg * (1.0 - t))
    ts = t * (1.0 - eps) + eps / C if eps > 0 else t
    loss = -(ts * lp * w).sum(-1)
    return loss.mean() if reduction == "mean" else loss


def supcon_loss(features: Tensor, labels: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                temperature: float = 0.07, base_temperature: float = 0.07, contrast_mode: str = "all") -> Tensor:
    """SupConLoss.forward (models/SupConLoss/loss.py:21-98).  features [bsz, n_views, D]."""
    if features.dim() < 3:
        raise ValueError("features needs to be [bsz, n_views, ...]")
    features = features.reshape(features.shape[0], features.shape[1], -1)
    bsz, nv, _ = features.shape
    if labels is not None and mask is not None:
        raise ValueError("Cannot define both labels and mask")
    if labels is None and mask is None:
        mask = torch.eye(bsz, dtype=torch.float32)
    elif labels is not None:
        labels = labels.reshape(-1, 1)
        if labels.shape[0] != bsz:
            raise ValueError("Num of labels does not match num of features")
        mask = (labels == labels.t()).float()
    else:
        mask = mask.float()            # the reference casts the mask to fp32 (loss.py:55)
    contrast = torch.cat([features[:, v] for v in range(nv)], dim=0)  # view-major (:58)
    if contrast_mode == "one":
        anchor, na = features[:, 0], 1
    elif contrast_mode == "all":
        anchor, na = contrast, nv
    else:
        raise ValueError(contrast_mode)
    a = anchor @ contrast.t() / temperature
    a = a - a.max(dim=1, keepdim=True).values.detach()
    mask = mask.repeat(na, nv)
    not_self = 1.0 - torch.eye(bsz * na, bsz * nv, dtype=torch.float32)
    mask = mask * not_self
    log_prob = a - torch.log((torch.exp(a) * not_self).sum(1, keepdim=True))
    mlpp = (mask * log_prob).sum(1) / mask.sum(1)
    return (-(temperature / base_temperature) * mlpp).view(na, bsz).mean()


def mlm_nll(logits: Tensor, target: Tensor) -> Tensor:
    """NLLLoss()(log_softmax(logits,-1).permute(0,2,1), target): mean over ALL B*T positions,
    target 0 is a real class (pretrain/roco_utils.py:235-236)."""
    lp = torch.log_softmax(logits, dim=-1)
    return -lp.gather(-1, target.long().unsqueeze(-1)).mean()


# ----------------------------------------------------------------------------------
# caption-similarity mask  (models/SupConLoss/supcon_utils.py:110-138)
# ----------------------------------------------------------------------------------
def jaccard_similarity(doc1: str, doc2: str) -> float:
    """SimilarityCalculator.jaccard_similarity, supcon_utils.py:120-138: |words1 & words2| / |words1 | words2| over the
    lower-cased whitespace-split word sets; 0.0 when both documents are empty."""
    w1, w2 = set(doc1.lower().split()), set(doc2.lower().split())
    union = w1 | w2
    if len(union) == 0:
        return 0.0
    return float(len(w1 & w2)) / len(union)


def jaccard_mask(caption, aug, bsz: int) -> Tensor:
    """SimilarityCalculator.jaccard, supcon_utils.py:110-118: [bsz, bsz] float32, ones on the diagonal."""
    mask = torch.zeros(bsz, bsz, dtype=torch.float)
    for c1 in range(len(caption)):
        for c2 in range(len(aug)):
            mask[c1, c2] = jaccard_similarity(caption[c1], aug[c2]) if c1 != c2 else 1.0
    return mask
