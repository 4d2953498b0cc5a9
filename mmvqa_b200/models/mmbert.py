"""MMBERT model surface.  Mirrors models/mmbert.py of the reference: get_bert_model,
get_transformer_model, TransformerAbstract, Transformer, RealFormer, Model, mean_pooling -- same
constructors, forward signatures, attribute names and state-dict keys, so vqamed2019/train.py and
pretrain/roco_*.py can import this package as ``models`` unchanged."""
import torch
import torch.nn as nn
from transformers import AutoModel

from .. import functional as Fn
from .._lib import ACT_SERF
from ..config import compute_dtype
from .image_encoding import get_transfer
from .realformer import ResEncoderBlock, run_blocks
from .serf import SERF
from .transformer import BertLayer


def _seed():
    return int(torch.randint(0, 2 ** 31 - 1, (1,)).item())


def get_bert_model(args):
    if args.task == 'distillation':
        bert_name = args.clinicalbert
    else:
        bert_name = 'bert-base-uncased'
    return bert_name


def get_transformer_model(args):
    if 'feedback-transformer' in args.transformer_model:
        raise NotImplementedError("FeedbackTransformer is outside the fusion-encoder hot path (SURVEY.md section 2, #11)")
    elif 'realformer' in args.transformer_model:
        return RealFormer(args)
    elif 'transformer' in args.transformer_model:
        return Transformer(args)
    else:
        raise NotImplementedError


class TransformerAbstract(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.bert_embedding = self.get_bert_embedding(args)
        self.trans = get_transfer(args)

    def get_bert_embedding(self, args):
        bert_name = get_bert_model(args)
        base_model = AutoModel.from_pretrained(bert_name)
        bert_model = nn.Sequential(*list(base_model.children())[0:])
        return bert_model[0]                      # HF BertEmbeddings: parameter holder for the fused kernel

    def prepare_input(self, img, input_ids, token_type_ids, mask):
        """mmbert.py:60-67: embed the tokens, then overwrite positions 0..num_vis-1 of every sample with
        the visual tokens.  One fused kernel (gather + LayerNorm + dropout + scatter) instead of HF
        BertEmbeddings followed by B*num_vis Python-level copies."""
        vizs = list(self.trans(img))
        return self.fuse(vizs, input_ids, token_type_ids)

    def fuse(self, vizs, input_ids, token_type_ids):
        emb = self.bert_embedding
        if torch.is_tensor(vizs):
            vis = vizs                                                                        # already [nvis, B, H]
        else:
            vis = torch.stack([v.float() for v in vizs], dim=0) if len(vizs) > 0 else None   # [nvis, B, H]
        p = emb.dropout.p if self.training else 0.0
        pad = emb.word_embeddings.padding_idx
        return Fn.EmbedFuseFn.apply(input_ids, token_type_ids, emb.word_embeddings.weight,
                                    emb.position_embeddings.weight,
                                    emb.token_type_embeddings.weight, emb.LayerNorm.weight, emb.LayerNorm.bias, vis,
                                    float(emb.LayerNorm.eps), p, _seed() if p > 0 else 0, compute_dtype(),
                                    -1 if pad is None else int(pad))


class Transformer(TransformerAbstract):
    def __init__(self, args):
        super().__init__(args)
        self.blocks = BertLayer(args, share='none', norm='pre')
        self.n_layers = args.n_layers

    def encode(self, h, mask):
        for i in range(self.n_layers):
            h = self.blocks(h, mask, i)
        return h

    def forward(self, img, input_ids, token_type_ids, mask):
        h = self.prepare_input(img, input_ids, token_type_ids, mask)
        return self.encode(h, mask)


class RealFormer(TransformerAbstract):
    def __init__(self, args):
        super().__init__(args)
        head_cnt = 8     # hard-coded in the reference (mmbert.py:100)
        self.mains = nn.Sequential(*[ResEncoderBlock(emb_s=args.hidden_size // head_cnt, head_cnt=head_cnt, dp1=0.1,
                                                     dp2=0.1) for _ in range(args.n_layers)])

    def encode(self, h, mask):
        """All blocks in one autograd node; `prev` is threaded inside and the last one dropped (mmbert.py:105-108)."""
        h, _ = run_blocks(list(self.mains), h, None, mask, self.training)
        return h

    def forward(self, img, input_ids, token_type_ids, mask):
        h = self.prepare_input(img, input_ids, token_type_ids, mask)
        return self.encode(h, mask)


class Model(nn.Module):
    def __init__(self, args, feat_dim=128):
        super(Model, self).__init__()
        self.transformer = get_transformer_model(args)
        self.fc1 = nn.Linear(args.hidden_size, args.hidden_size)
        self.activ1 = SERF()
        self.classifier = nn.Sequential(nn.Linear(args.hidden_size, args.hidden_size),
                                        nn.LayerNorm(args.hidden_size, eps=1e-12, elementwise_affine=True),
                                        nn.Linear(args.hidden_size, args.vocab_size))
        self.task = args.task
        self.dataset = args.dataset
        self.supcon = args.supcon if hasattr(args, 'supcon') else False
        if self.supcon:
            self.head = nn.Sequential(
                nn.Linear(args.hidden_size, args.hidden_size),
                SERF(),
                nn.Linear(args.hidden_size, feat_dim)
            )

    # classifier(SERF(fc1(z))): fc1+bias+SERF is one GEMM; Linear -> LN(1e-12) -> Linear; fp32 logits
    def _classify(self, z):
        z = Fn.linear(z, self.fc1.weight, self.fc1.bias, act=ACT_SERF)
        c0, ln, c2 = self.classifier[0], self.classifier[1], self.classifier[2]
        z = Fn.linear(z, c0.weight, c0.bias)
        z = Fn.add_layer_norm(z, None, ln.weight, ln.bias, ln.eps)
        return Fn.linear(z, c2.weight, c2.bias, out_fp32=True)

    def mlm_loss(self, h, target, chunk=4096):
        """mean over ALL B*T positions of NLL(log_softmax(classifier(SERF(fc1(h))))) -- the MLM objective of
        pretrain/roco_utils.py:235-236 on top of mmbert.py:154-155 -- with the vocabulary projection fused into a chunked
        cross entropy (Fn.chunked_vocab_ce): same value and gradients as
        ``NLLLoss()(log_softmax(self._classify(h), -1).permute(0, 2, 1), target)`` without the [B, T, V] logits.
        h = encoder output [B, T, hidden] (``model.encode_features`` / ``model.transformer``)."""
        z = Fn.linear(h, self.fc1.weight, self.fc1.bias, act=ACT_SERF)
        c0, ln, c2 = self.classifier[0], self.classifier[1], self.classifier[2]
        z = Fn.linear(z, c0.weight, c0.bias)
        z = Fn.add_layer_norm(z, None, ln.weight, ln.bias, ln.eps)
        rows = Fn.chunked_vocab_ce(z.reshape(-1, z.shape[-1]), c2.weight, c2.bias, target.reshape(-1), chunk)
        return rows.mean()

    def encode_features(self, feats, input_ids, segment_ids, input_mask):
        """feature maps -> encoder output [B, T, hidden] (the input of heads() / mlm_loss())."""
        tr = self.transformer
        h = tr.fuse(tr.trans.project_stacked(feats), input_ids, segment_ids)
        return tr.encode(h, input_mask)

    def _project(self, z):
        z = Fn.linear(z, self.head[0].weight, self.head[0].bias, act=ACT_SERF)
        z = Fn.linear(z, self.head[2].weight, self.head[2].bias, out_fp32=True)
        return Fn.L2NormFn.apply(z)

    def heads(self, h, input_mask):
        """Everything after the encoder (mmbert.py:151-167)."""
        if self.dataset == 'roco':
            if self.task == 'MLM':
                logits = self._classify(h)
                if self.supcon:
                    feat = self._project(mean_pooling(h, input_mask))
                    return logits, feat
            elif self.task == 'distillation':
                logits = h
            return logits
        elif self.dataset == 'VQA-Med':
            logits = self._classify(mean_pooling(h, input_mask))
            return logits, 0, 0

    def forward(self, img, input_ids, segment_ids, input_mask):
        h = self.transformer(img, input_ids, segment_ids, input_mask)
        return self.heads(h, input_mask)

    def forward_features(self, feats, input_ids, segment_ids, input_mask):
        """Hot path only: backbone feature maps (token order) -> outputs.  Same arithmetic as forward()
        after the CNN; used by bench.py and the parity tests, whose inputs start at the feature maps."""
        tr = self.transformer
        h = tr.fuse(tr.trans.project_stacked(feats), input_ids, segment_ids)
        return self.heads(tr.encode(h, input_mask), input_mask)


def mean_pooling(token_embeddings, attention_mask):
    """mask-weighted mean over tokens, denominator clamped at 1e-9 (mmbert.py:169-172)."""
    return Fn.MaskedMeanFn.apply(Fn.to_compute(token_embeddings), attention_mask.to(torch.float32).contiguous())
