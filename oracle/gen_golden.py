"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference; the GPU box does not have it).
The reference modules are imported as they are; three things they pull from the network
or from missing packages are stubbed *around* them:

  * ``timm``            -> fake module whose create_model() returns a tiny feature-list CNN
  * ``AutoModel.from_pretrained`` -> BertModel(BertConfig(small)) (random init, offline)
  * ``torchvision.models.resnet152(pretrained=True)`` -> tiny 10-child stand-in with the
    reference's channel table [64, 256, 512, 1024, 2048]

Usage:  python oracle/gen_golden.py          (writes tests/golden/*.pt)
"""
import contextlib
import io
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get("MMVQA_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


# ------------------------------------------------------------------ stubs
class TinyFeatureCNN(nn.Module):
    """Stand-in for timm tf_efficientnetv2_m(features_only=True): 5 maps, channels 24/48/80/176/512."""
    chans = [24, 48, 80, 176, 512]

    def __init__(self):
        super().__init__()
        cin, layers = 3, []
        for c, s in zip(self.chans, [2, 2, 1, 2, 2]):
            layers.append(nn.Sequential(nn.Conv2d(cin, c, 3, stride=s, padding=1), nn.SiLU()))
            cin = c
        self.stages = nn.ModuleList(layers)

    def forward(self, x):
        outs = []
        for st in self.stages:
            x = st(x)
            outs.append(x)
        return outs


class TinyResNet(nn.Module):
    """10 children like torchvision resnet: conv1,bn1,relu,maxpool,layer1..4,avgpool,fc."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 3, stride=2, padding=1)
        self.bn1 = nn.Identity()
        self.relu = nn.ReLU()
        self.maxpool = nn.MaxPool2d(2)
        self.layer1 = nn.Sequential(nn.Conv2d(64, 256, 1), nn.ReLU())
        self.layer2 = nn.Sequential(nn.Conv2d(256, 512, 1, stride=2), nn.ReLU())
        self.layer3 = nn.Sequential(nn.Conv2d(512, 1024, 1, stride=2), nn.ReLU())
        self.layer4 = nn.Sequential(nn.Conv2d(1024, 2048, 1, stride=2), nn.ReLU())
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(2048, 10)


class FixedFeatures(nn.Module):
    """Backbone stand-in that hands back pre-computed feature maps (so they can carry .grad)."""

    def __init__(self, feats):
        super().__init__()
        self._feats = feats

    def forward(self, img):
        return self._feats


def install_stubs(hidden, vocab, max_pos):
    import importlib.machinery
    import transformers  # noqa: F401  (must be imported before the fake timm is injected)
    from transformers import BertConfig, BertModel
    fake = types.ModuleType("timm")
    fake.__spec__ = importlib.machinery.ModuleSpec("timm", None)
    fake.create_model = lambda name, features_only=True, pretrained=True: TinyFeatureCNN()
    sys.modules["timm"] = fake
    transformers.AutoModel.from_pretrained = staticmethod(
        lambda name, *a, **k: BertModel(BertConfig(hidden_size=hidden, vocab_size=vocab, num_hidden_layers=1,
                                                   num_attention_heads=4, intermediate_size=hidden * 2,
                                                   max_position_embeddings=max_pos)))
    if REF not in sys.path:
        sys.path.insert(0, REF)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def sd(mod):
    return {k: v.detach().clone() for k, v in mod.state_dict().items()}


def grads(mod):
    return {k: p.grad.detach().clone() for k, p in mod.named_parameters() if p.grad is not None}


def save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".pt")
    torch.save(obj, path)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


# ------------------------------------------------------------------ cases
def case_activations():
    from models.serf import SERF
    from models.transformer import gelu
    with quiet():
        act = SERF()
    x = torch.tensor([-100.0, -30.0, -10.0, -3.0, -1.0, -0.25, 0.0, 0.25, 0.5, 1.0, 3.0, 10.0, 49.0, 50.0, 51.0,
                      80.0, 1000.0], dtype=torch.float64, requires_grad=True)
    y = act(x)
    (g,) = torch.autograd.grad(y.sum(), x)
    torch.manual_seed(1)
    xr = (torch.randn(4, 257) * 3).requires_grad_(True)
    yr = act(xr)
    (gr,) = torch.autograd.grad(yr, xr, torch.ones_like(yr))
    xg = torch.tensor([-4.0, -1.0, 0.0, 1.0, 2.0, 5.0], dtype=torch.float64, requires_grad=True)
    yg = gelu(xg)
    (gg,) = torch.autograd.grad(yg.sum(), xg)
    save("activations", {"serf_x64": x.detach(), "serf_y64": y.detach(), "serf_g64": g,
                         "serf_x32": xr.detach(), "serf_y32": yr.detach(), "serf_g32": gr,
                         "gelu_x64": xg.detach(), "gelu_y64": yg.detach(), "gelu_g64": gg})


def case_losses():
    from models.asl_singlelabel import ASLSingleLabel
    from models.SupConLoss.loss import SupConLoss
    out = {}
    # known-answer case quoted in SURVEY.md section 4
    lg = torch.tensor([[2.0, 0.5, -1.0, 0.0], [0.1, 0.2, 0.3, 0.4]], dtype=torch.float64, requires_grad=True)
    tg = torch.tensor([0, 3])
    crit = ASLSingleLabel()
    l = crit(lg, tg)
    (g,) = torch.autograd.grad(l, lg)
    out["asl_kat"] = {"logits": lg.detach(), "target": tg, "loss": l.detach(), "grad": g,
                      "targets_classes": crit.targets_classes.detach()}
    torch.manual_seed(2)
    for tag, kw in {"default": {}, "g1_2_eps0": dict(gamma_pos=1, gamma_neg=2, eps=0.0),
                    "sum": dict(reduction="none")}.items():
        lg = (torch.randn(9, 53) * 2).requires_grad_(True)
        tg = torch.randint(0, 53, (9,))
        l = ASLSingleLabel(**kw)(lg, tg)
        (g,) = torch.autograd.grad(l.sum(), lg)
        out["asl_" + tag] = {"kw": kw, "logits": lg.detach(), "target": tg, "loss": l.detach(), "grad": g}
    # SupCon
    f = torch.tensor([[[1, 0, 0], [.8, .6, 0]], [[0, 1, 0], [0, .6, .8]], [[0, 0, 1], [.6, 0, .8]]], dtype=torch.float64)
    f = f / f.norm(dim=-1, keepdim=True)
    soft = torch.tensor([[1, .5, 0], [.25, 1, .1], [0, .3, 1]], dtype=torch.float64)
    crit = SupConLoss()
    out["supcon_kat"] = {"features": f, "soft": soft, "labels": torch.tensor([0, 0, 1]),
                         "simclr": crit(f), "soft_loss": crit(f, mask=soft),
                         "label_loss": crit(f, labels=torch.tensor([0, 0, 1]))}
    torch.manual_seed(3)
    f = torch.randn(12, 2, 32)
    f = (f / f.norm(dim=-1, keepdim=True)).requires_grad_(True)
    soft = torch.rand(12, 12)
    soft.fill_diagonal_(1.0)
    labels = torch.randint(0, 4, (12,))
    rec = {"features": f.detach(), "soft": soft, "labels": labels}
    for tag, kw in {"simclr": {}, "soft": {"mask": soft}, "labels": {"labels": labels}}.items():
        l = crit(f, **kw)
        (g,) = torch.autograd.grad(l, f)
        rec[tag + "_loss"], rec[tag + "_grad"] = l.detach(), g
    l = SupConLoss(temperature=0.1, contrast_mode="one", base_temperature=0.07)(f, mask=soft)
    (g,) = torch.autograd.grad(l, f)
    rec["one_loss"], rec["one_grad"] = l.detach(), g
    out["supcon_rand"] = rec
    save("losses", out)


def case_transformer():
    from models.transformer import BertLayer, MultiHeadedSelfAttention
    out = {}
    args = types.SimpleNamespace(hidden_size=64, heads=4, hidden_dropout_prob=0.0, n_layers=2)
    torch.manual_seed(4)
    att = MultiHeadedSelfAttention(args)
    x = torch.randn(3, 9, 64, requires_grad=True)
    mask = torch.ones(3, 9, dtype=torch.long)
    mask[0, 6:] = 0
    mask[2, 3:] = 0
    y = att(x, mask)
    go = torch.randn_like(y)
    y.backward(go)
    out["mhsa"] = {"state": sd(att), "x": x.detach(), "mask": mask, "y": y.detach(), "scores": att.scores.detach(),
                   "go": go, "gx": x.grad.clone(), "gparams": grads(att)}
    args = types.SimpleNamespace(hidden_size=32, heads=4, hidden_dropout_prob=0.0, n_layers=2)
    for share, norm in [("none", "pre"), ("all", "post"), ("ffn", "pre"), ("att", "post"), ("none", "post"),
                        ("all", "pre")]:
        torch.manual_seed(5)
        layer = BertLayer(args, share=share, norm=norm)
        with torch.no_grad():   # make the LayerNorm affine non-trivial
            for n, p_ in layer.named_parameters():
                if "norm" in n:
                    p_.add_(0.1 * torch.randn_like(p_))
        x = torch.randn(3, 9, 32, requires_grad=True)
        h = x
        for i in range(args.n_layers):
            h = layer(h, mask, i)
        go = torch.randn_like(h)
        h.backward(go)
        out[f"bertlayer_{share}_{norm}"] = {"state": sd(layer), "x": x.detach(), "mask": mask, "y": h.detach(),
                                            "go": go, "gx": x.grad.clone(), "gparams": grads(layer)}
    save("transformer", out)


def case_realformer():
    from models.realformer import ResEncoderBlock
    torch.manual_seed(6)
    with quiet():
        blocks = nn.ModuleList([ResEncoderBlock(emb_s=8, head_cnt=8, dp1=0.0, dp2=0.0) for _ in range(3)])
    with torch.no_grad():
        for n, p_ in blocks.named_parameters():
            if "ln" in n:
                p_.add_(0.1 * torch.randn_like(p_))
    x = torch.randn(3, 10, 64, requires_grad=True)
    mask = torch.ones(3, 10, dtype=torch.long)
    mask[1, 7:] = 0
    mask[2, 2:] = 0
    h, prev, prevs = x, None, []
    for b in blocks:
        h, prev = b(h, prev=prev, mask=mask)
        prevs.append(prev.detach().clone())
    go = torch.randn_like(h)
    gp = torch.randn_like(prev) * 0.1
    (h * go).sum().add((prev * gp).sum()).backward()
    rec = {"state": sd(blocks), "x": x.detach(), "mask": mask, "y": h.detach(), "prevs": prevs, "go": go, "gprev": gp,
           "gx": x.grad.clone(), "gparams": grads(blocks)}
    # no-mask / no-prev single block
    torch.manual_seed(7)
    with quiet():
        b = ResEncoderBlock(emb_s=4, head_cnt=8, dp1=0.0, dp2=0.0)
    x2 = torch.randn(2, 5, 32)
    y2, p2 = b(x2)
    save("realformer", {"chain": rec, "single": {"state": sd(b), "x": x2, "y": y2.detach(), "prev": p2.detach()}})


def _vqa_inputs(B, T, vocab, num_vis=5):
    """Synthetic VQA-Med shaped tokens: [CLS], 5x0, [SEP], question, [SEP], pad (vqamed2019/utils.py:156-170)."""
    ids = torch.zeros(B, T, dtype=torch.long)
    seg = torch.zeros(B, T, dtype=torch.long)
    msk = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):
        qlen = int(torch.randint(2, T - 8, (1,)))
        ids[b, 0] = 5
        ids[b, 6] = 6
        ids[b, 7:7 + qlen] = torch.randint(7, vocab, (qlen,))
        ids[b, 7 + qlen] = 6
        seg[b, 7:8 + qlen] = 1
        msk[b, :8 + qlen] = 1
    return ids, seg, msk


def case_models():
    hidden, vocab, T, B = 64, 120, 20, 3
    install_stubs(hidden, vocab, 32)
    import models.image_encoding as ie
    import models.mmbert as mm
    ie.models_dict[5]["resnet152"][0] = lambda pretrained=True: TinyResNet()
    out = {}
    base = dict(task="MLM", clinicalbert="", num_vis=5, hidden_size=hidden, use_relu=False, heads=4,
                hidden_dropout_prob=0.0, n_layers=2, vocab_size=vocab)
    variants = {
        "vqa_realformer_effnet": dict(transformer_model="realformer", cnn_encoder="tf_efficientnetv2_m",
                                      dataset="VQA-Med", task="VQA"),
        "vqa_transformer_resnet_relu": dict(transformer_model="transformer", cnn_encoder="resnet152",
                                            dataset="VQA-Med", task="VQA", use_relu=True),
        "mlm_realformer_supcon": dict(transformer_model="realformer", cnn_encoder="tf_efficientnetv2_m",
                                      dataset="roco", task="MLM", supcon=True),
        "mlm_transformer": dict(transformer_model="transformer", cnn_encoder="tf_efficientnetv2_m",
                                dataset="roco", task="MLM"),
    }
    for seed, (name, over) in enumerate(variants.items(), start=10):
        args = types.SimpleNamespace(**{**base, **over})
        torch.manual_seed(seed)
        with quiet():
            model = mm.Model(args, feat_dim=16)
        model.eval()   # dropout off (RealFormer hard-codes p=0.1; BertEmbeddings p=0.1)
        with torch.no_grad():
            for n, p_ in model.named_parameters():
                if "LayerNorm" in n or ".ln" in n or "norm" in n or n.startswith("classifier.1"):
                    p_.add_(0.1 * torch.randn_like(p_))
        img = torch.randn(B, 3, 32, 32)
        ids, seg, msk = _vqa_inputs(B, T, vocab)
        trans = model.transformer.trans
        # capture the backbone feature maps (= the hot path's input) in the order the projector consumes them
        if "resnet" in args.cnn_encoder:
            ch = list(trans.model.children())
            feats = [nn.Sequential(*ch[:-k])(img) for k in (2, 3, 4, 5, 7)]
        else:
            feats = trans.model(img)
        feats = [f.detach() for f in feats]
        res = model(img, ids, seg, msk)
        state = {k: v for k, v in sd(model).items() if not k.startswith("transformer.trans.model.")}
        rec = {"args": vars(args), "state": state, "feats": feats, "ids": ids, "seg": seg, "mask": msk}
        if args.dataset == "VQA-Med":
            logits = res[0]
            assert res[1] == 0 and res[2] == 0
            tgt = torch.randint(0, vocab, (B,))
            from models.asl_singlelabel import ASLSingleLabel
            loss = ASLSingleLabel()(logits, tgt)
        elif getattr(args, "supcon", False):
            logits, feat = res
            rec["feat"] = feat.detach()
            tgt = torch.where(torch.rand(B, T) < 0.15, ids, torch.zeros_like(ids))
            loss = nn.NLLLoss()(logits.log_softmax(-1).permute(0, 2, 1), tgt) + (feat * torch.linspace(
                -1, 1, feat.numel()).view_as(feat)).sum()
        else:
            logits = res
            tgt = torch.where(torch.rand(B, T) < 0.15, ids, torch.zeros_like(ids))
            loss = nn.NLLLoss()(logits.log_softmax(-1).permute(0, 2, 1), tgt)
        loss.backward()
        g = grads(model)
        rec.update({"logits": logits.detach(), "target": tgt, "loss": loss.detach(),
                    "gparams": {k: v for k, v in g.items() if not k.startswith("transformer.trans.model.")}})
        out[name] = rec
    # projector alone, both activations, from the real Transfer classes
    for act_relu in (False, True):
        args = types.SimpleNamespace(**{**base, "cnn_encoder": "tf_efficientnetv2_m", "use_relu": act_relu})
        torch.manual_seed(20)
        with quiet():
            tr = ie.get_transfer(args)
        img = torch.randn(2, 3, 32, 32)
        feats = [f.detach().requires_grad_(True) for f in tr.model(img)]
        tr.model = FixedFeatures(feats)
        vs = tr(img)
        go = [torch.randn_like(v) for v in vs]
        sum((v * g_).sum() for v, g_ in zip(vs, go)).backward()
        out["projector_relu" if act_relu else "projector_serf"] = {
            "convs": [getattr(tr, n).weight.detach().clone() for n in ("conv2", "conv3", "conv4", "conv5", "conv7")],
            "gconvs": [getattr(tr, n).weight.grad.clone() for n in ("conv2", "conv3", "conv4", "conv5", "conv7")],
            "feats": [f.detach() for f in feats], "gfeats": [f.grad.clone() for f in feats],
            "vis": [v.detach() for v in vs], "go": go}
    save("models", out)


def case_jaccard():
    """SimilarityCalculator.jaccard (models/SupConLoss/supcon_utils.py:110-138).  The module itself cannot be imported
    here (sentence_transformers / bert_score / googletrans are missing), so the two methods are lifted out of the
    reference file with ``ast`` at generation time and executed unmodified on seeded captions."""
    import ast
    import random
    path = os.path.join(REF, "models", "SupConLoss", "supcon_utils.py")
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SimilarityCalculator")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("jaccard", "jaccard_similarity")]
    assert len(fns) == 2
    holder = ast.ClassDef(name="RefJaccard", bases=[], keywords=[], body=fns, decorator_list=[], type_params=[])
    mod = ast.Module(body=[holder], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"torch": torch}
    exec(compile(mod, path, "exec"), ns)
    ref = ns["RefJaccard"]()
    rng = random.Random(0)
    words = ["chest", "x-ray", "CT", "scan", "of", "the", "left", "right", "lung", "showing", "a", "Nodule", "mass", "MRI",
             "brain", "axial", "view", "with", "contrast", "no", "fracture", "Pleural", "effusion", "normal", "heart"]
    out = {}
    for name, bsz in (("small", 6), ("medium", 33)):
        cap = [" ".join(rng.choice(words) for _ in range(rng.randint(0 if name == "small" else 1, 14))) for _ in range(bsz)]
        aug = [" ".join(rng.choice(words) for _ in range(rng.randint(1, 14))) for _ in range(bsz)]
        if name == "small":
            cap[0], aug[1] = "", ""                # empty documents: union == 0 -> 0.0 (supcon_utils.py:136-138)
            cap[2] = "Lung  LUNG\tlung mass"       # case folding, repeated words, mixed whitespace
        with quiet():
            mask = ref.jaccard(cap, aug, bsz)
        out[name] = {"caption": cap, "aug": aug, "mask": mask.clone()}
    save("jaccard", out)


if __name__ == "__main__":
    torch.set_num_threads(4)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    case_activations()
    case_losses()
    case_transformer()
    case_realformer()
    case_models()
    case_jaccard()
