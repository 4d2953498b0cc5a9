"""BASELINE.json configs at full size (GPU): config[1] (EffNetV2-M maps + RealFormer-12 + ASL, B=16, T=28) against
the CPU oracle directly (it finishes in seconds), plus size-independent properties from SURVEY.md section 4."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import mmvqa_b200
    from mmvqa_b200.models.asl_singlelabel import ASLSingleLabel
    from mmvqa_b200.models.realformer import ResEncoderBlock, run_blocks

import bench
from oracle import mmbert_oracle as O

DEV = "cuda"


@pytest.fixture(scope="module")
def c2():
    model = bench.build_model(seed=0).eval()
    feats, ids, seg, mask, target = bench.synth_batch(16, 99)
    p = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    logits = O.model_forward(feats, ids, seg, mask, p, encoder="realformer", n_layers=bench.LAYERS, dataset="VQA-Med")
    loss = O.asl_single_label(logits, target)
    names = ["transformer.mains.0.kqv.weight", "transformer.mains.11.ff.2.weight", "transformer.mains.5.ln1.weight",
             "transformer.trans.conv2.weight", "transformer.trans.conv7.weight", "fc1.weight", "classifier.2.bias",
             "transformer.bert_embedding.position_embeddings.weight", "transformer.mains.3.proj.weight"]
    grads = dict(zip(names, torch.autograd.grad(loss, [p[n] for n in names])))
    return model, (feats, ids, seg, mask, target), logits.detach(), loss.detach(), grads


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_config1_full_step_vs_oracle(c2, dt):
    model, (feats, ids, seg, mask, target), ref_logits, ref_loss, ref_grads = c2
    model = model.to(DEV)
    with mmvqa_b200.compute_dtype_scope(dt):
        model.zero_grad(set_to_none=True)
        logits, z1, z2 = model.forward_features([f.to(DEV) for f in feats], ids.to(DEV), seg.to(DEV), mask.to(DEV))
        loss = ASLSingleLabel()(logits, target.to(DEV))
        loss.backward()
    params = dict(model.named_parameters())
    scale = ref_logits.abs().max()
    err = (logits.cpu() - ref_logits).abs().max() / scale
    if dt == torch.float32:
        assert err < 1e-4, err                               # fp32 path: 1e-4 of the logit range
        assert torch.equal(logits.argmax(-1).cpu(), ref_logits.argmax(-1)), "fp32 argmax must equal the reference's"
        assert abs(loss.item() - ref_loss.item()) < 1e-4 * abs(ref_loss.item())
        gtol = 2e-3
    else:
        assert err < 5e-2, err                               # bf16 storage / fp32 accumulate: 5e-2 of the range
        assert abs(loss.item() - ref_loss.item()) < 3e-2 * abs(ref_loss.item())
        gtol = 0.15
    for n, g in ref_grads.items():
        got = params[n].grad.cpu()
        e = (got - g).abs().max() / g.abs().max().clamp_min(1e-12)
        assert e < gtol, (n, float(e))
    model.cpu()


def test_realformer_mask_is_softmax_invariant_fullsize():
    """the query-side mask only shifts whole softmax rows, so the block output does not depend on
    it, while prev carries -10000 * L on masked rows (L = 12 layers, B = 16, T = 28, hidden 768)."""
    with mmvqa_b200.compute_dtype_scope(torch.float32):
        torch.manual_seed(1)
        blocks = [ResEncoderBlock(emb_s=96, head_cnt=8, dp1=0.0, dp2=0.0).to(DEV) for _ in range(12)]
        x = torch.randn(16, 28, 768, device=DEV)
        mask = torch.ones(16, 28, dtype=torch.long, device=DEV)
        mask[:, 20:] = 0
        y0, p0 = run_blocks(blocks, x, None, None, False)
        y1, p1 = run_blocks(blocks, x, None, mask, False)
        # invariant up to fp32 rounding: scores of magnitude 1.2e5 have an ulp of 0.0078, so the masked rows'
        # softmax (and, through the attended padded keys, every later activation) moves by ~1e-2
        torch.testing.assert_close(y1, y0, rtol=0, atol=5e-2)
        torch.testing.assert_close(p1[:, :20], p0[:, :20], rtol=0, atol=0.2)
        torch.testing.assert_close(p1[:, 20:], p0[:, 20:] - 120000.0, rtol=0, atol=1.0)
        assert (p1[:, 20:] < -110000).all()


def test_dp_gradients_equal_big_batch_fullsize(c2):
    """averaged gradients of two half batches == gradients of the concatenated batch."""
    model, (feats, ids, seg, mask, target), *_ = c2
    model = model.to(DEV)
    crit = ASLSingleLabel()
    with mmvqa_b200.compute_dtype_scope(torch.float32):
        def grads(sl):
            model.zero_grad(set_to_none=True)
            lg, _, _ = model.forward_features([f[sl].to(DEV) for f in feats], ids[sl].to(DEV), seg[sl].to(DEV), mask[sl].to(DEV))
            crit(lg, target[sl].to(DEV)).backward()
            return {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        full = grads(slice(0, 16))
        a, b = grads(slice(0, 8)), grads(slice(8, 16))
    for n in ("transformer.mains.0.kqv.weight", "transformer.mains.11.ff.0.weight", "classifier.2.weight",
              "transformer.trans.conv4.weight"):
        avg = 0.5 * (a[n] + b[n])
        e = (avg - full[n]).abs().max() / full[n].abs().max()
        assert e < 1e-4, (n, float(e))
    model.cpu()


def _mlm_model(supcon):
    import types
    from transformers import BertConfig, BertModel
    from mmvqa_b200.models import image_encoding as IE
    from mmvqa_b200.models import mmbert as MM
    import torch.nn as nn
    args = types.SimpleNamespace(task="MLM", clinicalbert="", transformer_model="realformer", cnn_encoder="tf_efficientnetv2_m",
                                 num_vis=5, hidden_size=768, use_relu=False, heads=8, hidden_dropout_prob=0.1, n_layers=2,
                                 vocab_size=30522, dataset="roco", supcon=supcon)
    old = MM.AutoModel.from_pretrained
    MM.AutoModel.from_pretrained = staticmethod(lambda name, *a, **k: BertModel(BertConfig(num_hidden_layers=1)))
    IE.models_dict[5]["tf_efficientnetv2_m"][0] = lambda *a, **k: nn.Identity()
    try:
        torch.manual_seed(5)
        return MM.Model(args).eval()
    finally:
        MM.AutoModel.from_pretrained = old


def test_mlm_supcon_full_width_vs_oracle():
    """BASELINE configs[2]/[3] shapes at full width (hidden 768, T=75, V=30522, feat 128; 2 layers and B=8 so the CPU
    oracle finishes in seconds): MLM logits [B,75,30522], SupCon features, fused CE == log_softmax+NLL, SupCon loss."""
    from mmvqa_b200 import functional as Fn
    from mmvqa_b200.models.SupConLoss.loss import SupConLoss
    B, Tn, V = 8, 75, 30522
    model = _mlm_model(True)
    g = torch.Generator().manual_seed(11)
    feats = [torch.randn(B, c, s, s, generator=g).abs() for c, s in bench.EFFNET_MAPS]
    ids = torch.randint(1000, V, (B, Tn), generator=g)
    ids[:, :5] = 0
    seg = torch.zeros(B, Tn, dtype=torch.long)
    mask = torch.ones(B, Tn, dtype=torch.long)
    mask[:, 60:] = 0
    target = torch.where(torch.rand(B, Tn, generator=g) < 0.15, ids, torch.zeros_like(ids))
    p = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        ref_logits, ref_feat = O.model_forward(feats, ids, seg, mask, p, encoder="realformer", n_layers=2, dataset="roco",
                                               supcon=True)
        ref_mlm = O.mlm_nll(ref_logits, target)
        pairs = ref_feat.view(B // 2, 2, -1)
        ref_sc = O.supcon_loss(pairs)
    model = model.to(DEV)
    for dt, tol in ((torch.float32, 2e-4), (torch.bfloat16, 6e-2)):
        with mmvqa_b200.compute_dtype_scope(dt):
            logits, feat = model.forward_features([f.to(DEV) for f in feats], ids.to(DEV), seg.to(DEV), mask.to(DEV))
            assert logits.shape == (B, Tn, V) and logits.dtype == torch.float32 and feat.shape == (B, 128)
            err = (logits.cpu() - ref_logits).abs().max() / ref_logits.abs().max()
            assert err < tol, (dt, float(err))
            ferr = (feat.cpu() - ref_feat).abs().max()
            assert ferr < (1e-4 if dt == torch.float32 else 3e-2), (dt, float(ferr))
            rows = Fn.CrossEntropyRowsFn.apply(logits.view(B * Tn, V), target.view(-1).to(DEV))
            assert abs(rows.mean().item() - ref_mlm.item()) < (1e-4 if dt == torch.float32 else 3e-2) * ref_mlm.item()
            sc = SupConLoss()(feat.view(B // 2, 2, -1))
            assert abs(sc.item() - ref_sc.item()) < (1e-4 if dt == torch.float32 else 5e-2) * abs(ref_sc.item())
            if dt == torch.float32:
                assert torch.equal(logits.argmax(-1).cpu(), ref_logits.argmax(-1))
            (rows.mean() + sc).backward()
            gsum = sum(float(q.grad.abs().sum()) for q in model.parameters() if q.grad is not None)
            assert gsum > 0 and gsum == gsum
            model.zero_grad(set_to_none=True)


def test_supcon_global_batch_2048():
    """BASELINE configs[3]: N = 2048 contrast rows (1024 pairs), D = 128, soft jaccard-shaped mask, vs the oracle."""
    from mmvqa_b200.models.SupConLoss.loss import SupConLoss
    g = torch.Generator().manual_seed(12)
    f = torch.randn(1024, 2, 128, generator=g)
    f = f / f.norm(dim=-1, keepdim=True)
    soft = torch.rand(1024, 1024, generator=g)
    soft.fill_diagonal_(1.0)
    fl = f.clone().requires_grad_(True)
    ref = O.supcon_loss(fl, mask=soft)
    (gref,) = torch.autograd.grad(ref, fl)
    for dt, tol in ((torch.float32, 1e-4), (torch.bfloat16, 3e-2)):
        with mmvqa_b200.compute_dtype_scope(dt):
            fd = f.to(DEV).requires_grad_(True)
            loss = SupConLoss()(fd, mask=soft.to(DEV))
            loss.backward()
            assert abs(loss.item() - ref.item()) < tol * abs(ref.item()), (dt, loss.item(), ref.item())
            e = (fd.grad.cpu() - gref).abs().max() / gref.abs().max()
            assert e < (1e-3 if dt == torch.float32 else 0.15), (dt, float(e))


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[0] and configs[2] at their real sizes (only hidden-64 goldens covered them)
# ------------------------------------------------------------------------------------------------------------------
RESNET_MAPS = [(2048, 7), (1024, 14), (512, 28), (256, 56), (64, 112)]      # token order deep -> shallow (image_encoding.py:13,87)


def _full_model(transformer_model, cnn_encoder, heads, n_layers, dataset, task, supcon=False, seed=7):
    import types
    import torch.nn as nn
    from transformers import BertConfig, BertModel
    from mmvqa_b200.models import image_encoding as IE
    from mmvqa_b200.models import mmbert as MM
    args = types.SimpleNamespace(task=task, clinicalbert="", transformer_model=transformer_model, cnn_encoder=cnn_encoder,
                                 num_vis=5, hidden_size=768, use_relu=False, heads=heads, hidden_dropout_prob=0.1,
                                 n_layers=n_layers, vocab_size=30522, dataset=dataset, supcon=supcon)
    old = MM.AutoModel.from_pretrained
    MM.AutoModel.from_pretrained = staticmethod(lambda name, *a, **k: BertModel(BertConfig(num_hidden_layers=1)))
    IE.models_dict[5]["tf_efficientnetv2_m"][0] = lambda *a, **k: nn.Identity()
    IE.models_dict[5]["resnet152"][0] = lambda *a, **k: nn.Identity()
    try:
        torch.manual_seed(seed)
        return MM.Model(args).eval()
    finally:
        MM.AutoModel.from_pretrained = old


def test_c1_resnet152_transformer12_forward_fullsize():
    """configs[0]: ResNet152 map shapes (2048@7^2 ... 64@112^2, deep -> shallow token order), Transformer-12 (12 heads x
    64, shared pre-norm), hidden 768, num_vis 5, B = 16, T = 28, un-swapped vocab head -> logits [16, 30522], forward.
    fp32 path: 1e-4 of the logit range and bit-exact argmax; bf16 path: 5e-2 of the range."""
    B = 16
    model = _full_model("transformer", "resnet152", 12, 12, "VQA-Med", "VQA")
    g = torch.Generator().manual_seed(21)
    feats = [torch.randn(B, c, s, s, generator=g).abs() for c, s in RESNET_MAPS]
    _, ids, seg, mask, _ = bench.synth_batch(B, 5)
    p = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        ref = O.model_forward(feats, ids, seg, mask, p, encoder="transformer", n_layers=12, heads=12, dataset="VQA-Med")
    assert ref.shape == (B, 30522)
    model = model.to(DEV)
    for dt, tol in ((torch.float32, 1e-4), (torch.bfloat16, 5e-2)):
        with mmvqa_b200.compute_dtype_scope(dt), torch.no_grad():
            logits, z1, z2 = model.forward_features([f.to(DEV) for f in feats], ids.to(DEV), seg.to(DEV), mask.to(DEV))
        assert z1 == 0 and z2 == 0 and logits.shape == (B, 30522)
        err = ((logits.float().cpu() - ref).abs().max() / ref.abs().max()).item()
        assert err < tol, (dt, err)
        if dt == torch.float32:
            assert torch.equal(logits.argmax(-1).cpu(), ref.argmax(-1)), "fp32 argmax must equal the reference's"
    model.cpu()


def test_c3_mlm_step_at_the_per_gpu_shape():
    """configs[2] on one of its 8 GPUs: RealFormer-12 MLM step, B = 32 (256 / 8), T = 75, V = 30522, EffNetV2-M maps,
    forward + NLL over every position (roco_utils.py:235-236, target 0 is a class) + backward, against the oracle."""
    from mmvqa_b200 import functional as Fn
    B, Tn, V = 32, 75, 30522
    model = _full_model("realformer", "tf_efficientnetv2_m", 8, 12, "roco", "MLM")
    g = torch.Generator().manual_seed(31)
    feats = [torch.randn(B, c, s, s, generator=g).abs() for c, s in bench.EFFNET_MAPS]
    ids = torch.randint(1000, V, (B, Tn), generator=g)
    ids[:, :5] = 0
    seg = torch.zeros(B, Tn, dtype=torch.long)
    mask = torch.ones(B, Tn, dtype=torch.long)
    for b in range(B):
        mask[b, 40 + b:] = 0
    target = torch.where(torch.rand(B, Tn, generator=g) < 0.15, ids, torch.zeros_like(ids))
    names = ["transformer.mains.0.kqv.weight", "transformer.mains.11.ff.0.weight", "transformer.mains.6.ln2.bias",
             "classifier.2.weight", "fc1.bias", "transformer.trans.conv5.weight"]
    p = {k: v.detach().clone().requires_grad_(k in names) for k, v in model.state_dict().items()}
    ref_logits = O.model_forward(feats, ids, seg, mask, p, encoder="realformer", n_layers=12, dataset="roco")
    ref_loss = O.mlm_nll(ref_logits, target)
    ref_grads = dict(zip(names, torch.autograd.grad(ref_loss, [p[n] for n in names])))
    ref_logits = ref_logits.detach()
    model = model.to(DEV)
    params = dict(model.named_parameters())
    for dt, tol, gtol in ((torch.float32, 2e-4, 3e-3), (torch.bfloat16, 6e-2, 0.15)):
        with mmvqa_b200.compute_dtype_scope(dt):
            model.zero_grad(set_to_none=True)
            logits = model.forward_features([f.to(DEV) for f in feats], ids.to(DEV), seg.to(DEV), mask.to(DEV))
            assert logits.shape == (B, Tn, V)
            loss = Fn.CrossEntropyRowsFn.apply(logits.view(B * Tn, V), target.view(-1).to(DEV)).mean()
            loss.backward()
        err = ((logits.float().cpu() - ref_logits).abs().max() / ref_logits.abs().max()).item()
        assert err < tol, (dt, err)
        assert abs(loss.item() - ref_loss.item()) < (1e-4 if dt == torch.float32 else 3e-2) * ref_loss.item()
        if dt == torch.float32:
            # 2400 rows x 30522 classes of a random-init head: near-ties between the two largest logits exist at the
            # 1e-6 level, so the token-level argmax is checked as a rate (the VQA answer argmax above is bit-exact)
            same = (logits.argmax(-1).cpu() == ref_logits.argmax(-1)).float().mean().item()
            assert same >= 0.995, same
        for n, gr in ref_grads.items():
            e = ((params[n].grad.cpu() - gr).abs().max() / gr.abs().max().clamp_min(1e-12)).item()
            assert e < gtol, (dt, n, e)
        del logits, loss
    model.cpu()


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_chunked_vocab_ce_at_the_c3_shape(dt):
    """SURVEY section 8f-1 at M = 2400 (B = 32, T = 75), V = 30522, hidden 768: the chunked vocab GEMM + cross entropy
    (never allocates [M, V]) against O.mlm_nll on CPU -- loss, dh, dW, db -- and against the two-tensor path of the same
    library for peak memory."""
    from mmvqa_b200 import functional as Fn
    M, H, V = 2400, 768, 30522
    g = torch.Generator().manual_seed(41)
    h = torch.randn(M, H, generator=g)
    W = torch.randn(V, H, generator=g) * 0.05
    b = torch.randn(V, generator=g) * 0.1
    target = torch.where(torch.rand(M, generator=g) < 0.15, torch.randint(1000, V, (M,), generator=g), torch.zeros(M, dtype=torch.long))
    hr, Wr, br = h.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = O.mlm_nll((hr @ Wr.t() + br).view(1, M, V), target.view(1, M))
    ref.backward()
    with mmvqa_b200.compute_dtype_scope(dt):
        Fn.invalidate_weight_cache()
        hd = h.to(DEV).requires_grad_(True)
        Wd = torch.nn.Parameter(W.to(DEV))
        bd = torch.nn.Parameter(b.to(DEV))
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        loss = Fn.chunked_vocab_ce(hd, Wd, bd, target.to(DEV), 4096).mean()
        loss.backward()
        torch.cuda.synchronize()
        peak_chunked = torch.cuda.max_memory_allocated() - base
        tol, gtol = (1e-5, 2e-3) if dt == torch.float32 else (2e-3, 6e-2)
        assert abs(loss.item() - ref.item()) < tol * ref.item(), (loss.item(), ref.item())
        for got, want, nm in ((hd.grad, hr.grad, "dh"), (Wd.grad, Wr.grad, "dW"), (bd.grad, br.grad, "db")):
            e = ((got.float().cpu() - want).abs().max() / want.abs().max()).item()
            assert e < gtol, (nm, e)
        # the two-tensor path: fp32 logits [M, V] + their gradient
        hd2 = h.to(DEV).requires_grad_(True)
        Wd.grad = None
        bd.grad = None
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        logits = Fn.linear(hd2, Wd, bd, out_fp32=True)
        loss2 = Fn.CrossEntropyRowsFn.apply(logits, target.to(DEV)).mean()
        loss2.backward()
        torch.cuda.synchronize()
        peak_full = torch.cuda.max_memory_allocated() - base
        assert abs(loss2.item() - loss.item()) < (1e-5 if dt == torch.float32 else 2e-3) * abs(loss.item())
        assert peak_chunked < 0.6 * peak_full, (peak_chunked, peak_full)      # 2 x 293 MB of logits are gone
        Fn.invalidate_weight_cache()


def test_model_mlm_loss_equals_logits_path():
    """Model.mlm_loss (chunked) == NLL(log_softmax(Model logits)) on a 2-layer full-width model."""
    from mmvqa_b200 import functional as Fn
    B, Tn, V = 4, 75, 30522
    model = _mlm_model(False).to(DEV)
    g = torch.Generator().manual_seed(17)
    feats = [torch.randn(B, c, s, s, generator=g).abs().to(DEV) for c, s in bench.EFFNET_MAPS]
    ids = torch.randint(1000, V, (B, Tn), generator=g).to(DEV)
    seg = torch.zeros(B, Tn, dtype=torch.long, device=DEV)
    mask = torch.ones(B, Tn, dtype=torch.long, device=DEV)
    target = torch.where(torch.rand(B, Tn, generator=g) < 0.15, ids.cpu(), torch.zeros(B, Tn, dtype=torch.long)).to(DEV)
    for dt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 3e-3)):
        with mmvqa_b200.compute_dtype_scope(dt):
            Fn.invalidate_weight_cache()
            model.zero_grad(set_to_none=True)
            logits = model.forward_features(feats, ids, seg, mask)
            l_ref = Fn.CrossEntropyRowsFn.apply(logits.view(B * Tn, V), target.view(-1)).mean()
            l_ref.backward()
            g_ref = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
            model.zero_grad(set_to_none=True)
            l_new = model.mlm_loss(model.encode_features(feats, ids, seg, mask), target)
            l_new.backward()
            assert abs(l_new.item() - l_ref.item()) < tol * abs(l_ref.item())
            for n in ("classifier.2.weight", "classifier.2.bias", "fc1.weight", "transformer.mains.0.kqv.weight"):
                p = dict(model.named_parameters())[n]
                e = ((p.grad - g_ref[n]).abs().max() / g_ref[n].abs().max().clamp_min(1e-12)).item()
                assert e < (2e-3 if dt == torch.float32 else 8e-2), (dt, n, e)
    Fn.invalidate_weight_cache()
    model.cpu()
