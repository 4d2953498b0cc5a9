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
