"""CPU-side checks: the C-ABI library loads and exports every symbol include/mmvqa.h declares, the module
mirrors keep the reference's constructor surface / state-dict keys, and the product path fails loudly
(no CPU fallback).  No compute calls: there is no GPU here."""
import os
import re
import types

import pytest
import torch
import torch.nn as nn

import mmvqa_b200
from mmvqa_b200 import _lib
from mmvqa_b200.models import image_encoding as IE
from mmvqa_b200.models import mmbert as MM
from mmvqa_b200.models.asl_singlelabel import ASLSingleLabel
from mmvqa_b200.models.realformer import ResEncoderBlock
from mmvqa_b200.models.serf import SERF
from mmvqa_b200.models.SupConLoss.loss import SupConLoss
from mmvqa_b200.models.transformer import BertLayer, MultiHeadedSelfAttention, PositionWiseFeedForward

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "mmvqa.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(mmvqa_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 28
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = mmvqa_b200.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mmvqa_abi_version() == _lib.ABI_VERSION
    assert lib.mmvqa_launch_count() == 0


def test_gemm_args_struct_matches_header_order():
    hdr = open(os.path.join(ROOT, "include", "mmvqa.h")).read()
    body = hdr[hdr.index("typedef struct mmvqa_gemm_args {"):hdr.index("} mmvqa_gemm_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"(\w+)\s*$", part.strip())[0])
    assert names == [f[0] for f in _lib.GemmArgs._fields_]


def test_no_cpu_fallback():
    with pytest.raises(mmvqa_b200.MMVQAError):
        SERF()(torch.randn(4, 8))                      # CPU tensor -> loud failure, never torch math
    with pytest.raises(mmvqa_b200.MMVQAError):
        ASLSingleLabel()(torch.randn(2, 5), torch.tensor([0, 1]))


def _args(**over):
    base = dict(task="MLM", clinicalbert="", num_vis=5, hidden_size=64, use_relu=False, heads=4, hidden_dropout_prob=0.1,
                n_layers=2, vocab_size=120, transformer_model="realformer", cnn_encoder="tf_efficientnetv2_m",
                dataset="VQA-Med")
    base.update(over)
    return types.SimpleNamespace(**base)


def _model(**over):
    from transformers import BertConfig, BertModel
    old = MM.AutoModel.from_pretrained
    MM.AutoModel.from_pretrained = staticmethod(lambda name, *a, **k: BertModel(BertConfig(
        hidden_size=64, vocab_size=120, num_hidden_layers=1, num_attention_heads=4, intermediate_size=128,
        max_position_embeddings=32)))
    IE.models_dict[5]["tf_efficientnetv2_m"][0] = lambda *a, **k: nn.Identity()
    IE.models_dict[5]["resnet152"][0] = lambda *a, **k: nn.Identity()
    try:
        return MM.Model(_args(**over), feat_dim=16)
    finally:
        MM.AutoModel.from_pretrained = old


def test_state_dict_keys_match_the_reference(golden):
    g = golden("models")
    for name in ("vqa_realformer_effnet", "vqa_transformer_resnet_relu", "mlm_realformer_supcon", "mlm_transformer"):
        c = g[name]
        m = _model(**{k: v for k, v in c["args"].items()})
        ours = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith("transformer.trans.model.")}
        ref = {k: tuple(v.shape) for k, v in c["state"].items()}
        assert ours == ref, (set(ours) ^ set(ref))


def test_module_surface():
    a = _args()
    m = _model()
    assert isinstance(m.classifier, nn.Sequential) and isinstance(m.classifier[2], nn.Linear)
    m.classifier[2] = nn.Linear(64, 7)                 # vqamed2019/train.py:137 swaps the head
    assert hasattr(m.transformer, "bert_embedding") and hasattr(m.transformer, "trans") and hasattr(m.transformer, "mains")
    t = _model(transformer_model="transformer")
    assert isinstance(t.transformer.blocks, BertLayer) and t.transformer.n_layers == 2
    with pytest.raises(NotImplementedError):
        MM.get_transformer_model(_args(transformer_model="lstm"))
    with pytest.raises(NotImplementedError):
        IE.get_transfer(_args(cnn_encoder="vgg"))
    b = ResEncoderBlock(emb_s=8, head_cnt=8)
    assert tuple(b.kqv.weight.shape) == (24, 8) and b.kqv.bias is None and b.proj.bias is None
    assert [type(x).__name__ for x in b.ff] == ["Linear", "SERF", "Linear", "Dropout"]
    att = MultiHeadedSelfAttention(a)
    x = torch.zeros(2, 3, 64)
    assert att.split_last(x, (4, -1)).shape == (2, 3, 4, 16)
    assert att.merge_last(att.split_last(x, (4, -1)), 2).shape == (2, 3, 64)
    assert isinstance(PositionWiseFeedForward(a).fc1, nn.Linear)
    for share in ("all", "att", "ffn", "none"):
        BertLayer(a, share=share, norm="post")
    crit = ASLSingleLabel()
    assert (crit.gamma_pos, crit.gamma_neg, crit.eps, crit.reduction) == (0, 4, 0.1, "mean")
    s = SupConLoss()
    assert (s.temperature, s.contrast_mode, s.base_temperature) == (0.07, "all", 0.07)
    with pytest.raises(ValueError):
        s(torch.randn(4, 8))


def test_same_seed_gives_reference_parameter_order():
    """Parameter creation order follows the reference, so one torch.manual_seed gives the same init."""
    torch.manual_seed(6)
    ours = nn.ModuleList([ResEncoderBlock(emb_s=8, head_cnt=8, dp1=0.0, dp2=0.0) for _ in range(3)])
    names = [n for n, _ in ours.named_parameters()]
    assert names[:10] == ["0.kqv.weight", "0.proj.weight", "0.ln1.weight", "0.ln1.bias", "0.ln2.weight", "0.ln2.bias",
                          "0.ff.0.weight", "0.ff.0.bias", "0.ff.2.weight", "0.ff.2.bias"]


def test_compute_dtype_switch():
    assert mmvqa_b200.compute_dtype() in (torch.bfloat16, torch.float32)
    with mmvqa_b200.compute_dtype_scope("fp32"):
        assert mmvqa_b200.compute_dtype() == torch.float32
    with pytest.raises((ValueError, KeyError)):
        mmvqa_b200.set_compute_dtype("fp16")


def test_shim_aliases_reference_import_paths():
    import sys
    import mmvqa_b200.shim as shim
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    try:
        shim.install()
        from models.mmbert import Model, get_transformer_model, mean_pooling  # noqa: F401  (vqamed2019/train.py:21)
        from models.asl_singlelabel import ASLSingleLabel as A2               # noqa: F401  (train.py:22)
        from models.SupConLoss.loss import SupConLoss as S2                   # noqa: F401  (roco_supcon_train.py)
        from models.realformer import ResEncoderBlock as R2                   # noqa: F401
        from models.transformer import BertLayer as B2, gelu                  # noqa: F401
        from models.image_encoding import get_transfer, models_dict          # noqa: F401
        from models.serf import SERF as SF2                                   # noqa: F401
        assert Model is MM.Model and A2 is ASLSingleLabel and S2 is SupConLoss
    finally:
        shim.uninstall()
        sys.modules.update(saved)


def test_serf_hermite_table_accuracy():
    """The projector kernel evaluates SERF / SERF' from a cubic table of g(x) = erf(softplus(x)) (csrc/common.cuh).
    This replays the table construction and the lookup in numpy float32 -- same constants (parsed from the header), same
    formulas -- and checks them against the reference definition models/serf.py:23-24 in float64, including the
    saturated tails (x < -18: 0, x > 6: slope exactly 1, as the reference's clamp gives for x > 50)."""
    import math
    import re

    import numpy as np
    src = open(os.path.join(ROOT, "mmvqa_b200", "csrc", "common.cuh")).read()
    N = int(re.search(r"SERF_TAB_N = (\d+);", src).group(1))
    X0 = np.float32(re.search(r"SERF_TAB_X0 = (-?[\d.]+)f;", src).group(1))
    H = np.float32(re.search(r"SERF_TAB_H = ([\d.]+)f;", src).group(1))
    assert abs(float(X0) + N * float(H) - 6.0) < 1e-6

    def erf64(v):
        return np.vectorize(math.erf)(v)

    def node(i):
        if i <= 0:
            return np.float32(0), np.float32(0)
        if i >= N:
            return np.float32(1), np.float32(0)
        x = np.float64(X0 + H * np.float32(i))
        sp = np.log1p(np.exp(x))
        return np.float32(math.erf(sp)), np.float32(float(H) * 1.1283791670955126 * np.exp(-sp * sp) / (1.0 + np.exp(-x)))
    tab = np.zeros((N, 4), np.float32)
    for i in range(N):
        (g0, m0), (g1, m1) = node(i), node(i + 1)
        dl = np.float32(g1 - g0)
        tab[i] = [g0, m0, np.float32(3) * dl - np.float32(2) * m0 - m1, m0 + m1 - np.float32(2) * dl]
    x = np.concatenate([np.linspace(-40, 70, 400001), [-18.0, 6.0, 0.0, 50.0, 50.5]]).astype(np.float32)
    u = np.clip((x - X0) * np.float32(1.0 / H), 0, N - 0.001).astype(np.float32)
    i = u.astype(np.int32)
    t = (u - i).astype(np.float32)
    e = tab[i]
    g = e[:, 0] + t * (e[:, 1] + t * (e[:, 2] + t * e[:, 3]))
    gp = (e[:, 1] + t * (2 * e[:, 2] + 3 * t * e[:, 3])) / H
    a, d = x * g, g + x * gp
    xd = x.astype(np.float64)
    sp = np.log1p(np.exp(np.minimum(xd, 50.0)))
    a_ref = xd * erf64(sp)
    d_ref = np.where(xd > 50.0, erf64(sp), erf64(sp) + xd * 1.1283791670955126 * np.exp(-sp * sp) / (1.0 + np.exp(-xd)))
    assert np.abs(a - a_ref).max() < 2e-6, np.abs(a - a_ref).max()
    assert np.abs(d - d_ref).max() < 1e-5, np.abs(d - d_ref).max()
    assert np.all(a[x > 6.0] == x[x > 6.0]) and np.all(d[x > 6.5] == 1.0)      # SERF(x) = x, SERF'(x) = 1 in the saturated tail


def test_resnet_single_pass_taps_equal_the_five_prefix_runs():
    """SURVEY.md section 8 row f4 (backbone hand-off): the reference re-runs five prefixes of the ResNet from the image
    (image_encoding.py:72-85: children()[:-2], [:-3], [:-4], [:-5], [:-7]); ResNetTransfer here runs the backbone once
    and taps the same five tensors.  Bit-exact in eval mode (same modules, same inputs, same order of operations)."""
    from torchvision import models
    torch.manual_seed(0)
    backbone = models.resnet18(weights=None).eval()          # same children() layout as resnet152, CPU-sized
    img = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        taps = IE.ResNetTransfer.tap_feature_maps(backbone, img)
        ch = list(backbone.children())
        for tok, cut in enumerate((2, 3, 4, 5, 7)):
            want = nn.Sequential(*ch[:-cut])(img)
            assert taps[tok].shape == want.shape
            assert torch.equal(taps[tok], want), "token %d (prefix [:-%d]) differs" % (tok, cut)
    # deep -> shallow token order, the channel counts of models_dict scale with the architecture (resnet18: /4)
    assert [t.shape[1] for t in taps] == [512, 256, 128, 64, 64]
    assert [t.shape[-1] for t in taps] == [2, 4, 8, 16, 32]


def test_struct_layouts_of_round2_entry_points_match_the_header():
    """ctypes mirrors of the by-value / by-pointer structs added in round 2 (descriptor of the multi-tensor Adam with the
    row gate, prefetch and multi-cast lists) have the header's field order, and the numpy descriptor the optimizer uploads
    has the same size as the C struct."""
    import ctypes as C
    import numpy as np
    from mmvqa_b200 import optim
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "mmvqa.h")).read()
    body = re.search(r"typedef struct mmvqa_adam_desc \{(.*?)\} mmvqa_adam_desc;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"(\w+)\s*(?:;|,)", body)
    assert names == [f[0] for f in _lib.AdamDesc._fields_]
    assert C.sizeof(_lib.AdamDesc) == optim._DESC.itemsize == 72
    assert list(optim._DESC.names) == [f[0] for f in _lib.AdamDesc._fields_]
    assert int(re.search(r"#define MMVQA_PREFETCH_MAX (\d+)", hdr).group(1)) == _lib.PREFETCH_MAX
    assert int(re.search(r"#define MMVQA_CAST_MULTI_MAX (\d+)", hdr).group(1)) == _lib.CAST_MULTI_MAX
    assert [f[0] for f in _lib.PrefetchList._fields_] == ["ptr", "bytes", "n"]
    assert [f[0] for f in _lib.CastList._fields_] == ["src", "dst", "ld_src", "ld_dst", "rows", "cols", "src_bf16", "n"]
    assert int(re.search(r"#define MMVQA_ABI_VERSION (\d+)", hdr).group(1)) == _lib.ABI_VERSION


def test_sink_group_schedule_and_update_modes():
    """FusedAdam host logic (no kernels): a sink-group schedule (4, 2, 1) flushes after 4, then 2, then single layers and
    restarts at step(); hook groups bypass it; the update mode follows the size of the launch and the tail flag."""
    from mmvqa_b200 import optim
    ps = [nn.Parameter(torch.zeros(4, 4)) for _ in range(9)]
    opt = optim.FusedAdam(ps, lr=1e-3, sink_group=(4, 2, 1))
    flushed = []
    opt._flush_early = lambda gi, params, gs, side, background=True, fast=False: flushed.append((len(params), background, fast))
    for i, p in enumerate(ps[:8]):
        assert opt._sink([p], [torch.zeros(4, 4)], None, i == 7)
    assert [n for n, _, _ in flushed] == [4, 2, 1, 1]
    assert all(b for _, b, _ in flushed)                      # pending layers are bulk (background) updates
    assert opt._sink([ps[8]], [torch.zeros(4, 4)], None, from_hook=True)
    assert flushed[-1] == (1, False, False)                   # hook groups: critical-path mode, own stream
    opt._early_ids.clear()
    opt._sink_round = 0
    one = optim.FusedAdam([nn.Parameter(torch.zeros(2))], lr=1e-3)
    calls = []
    one._flush_early = lambda gi, params, gs, side, background=True, fast=False: calls.append((background, fast))
    one._sink(one.param_groups[0]["params"], [torch.zeros(2)], None, True)
    assert calls == [(True, True)]                            # last layer of the backward pass: fast bulk update
    assert optim.CHUNK < optim.CHUNK_BACKGROUND and optim.GATED_ROWS <= 1024
