"""One-launch RealFormer encoder forward (csrc/rf_encoder.cu, mmvqa_rf_encoder_fwd) against the per-operator chain of the
same library and against the CPU oracle (models/realformer.py:30-51 x n_layers, mmbert.py:103-108)."""
import os
import sys

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))

if torch.cuda.is_available():
    import mmvqa_b200
    from mmvqa_b200 import functional as Fn
    from mmvqa_b200 import ops
    from mmvqa_b200.models.realformer import ResEncoderBlock, block_params, run_blocks

NAMES = ["xin", "kqv", "scores", "attn", "y1", "mean1", "rstd1", "x1", "hpre", "hact", "y2", "mean2", "rstd2"]


def _blocks(n_layers, seed=0, drop=0.0):
    torch.manual_seed(seed)
    blocks = nn.ModuleList([ResEncoderBlock(emb_s=96, head_cnt=8, dp1=drop, dp2=drop) for _ in range(n_layers)]).cuda()
    for b in blocks:
        for p in (b.ln1.weight, b.ln2.weight):
            p.data.uniform_(0.5, 1.5)
        for p in (b.ln1.bias, b.ln2.bias, b.ff[0].bias, b.ff[2].bias):
            p.data.uniform_(-0.3, 0.3)
    return blocks


def _run(native, x, mask, prev, params, drop=0.0, seed=1234):
    os.environ["MMVQA_RF_ENCODER"] = "1" if native else "0"
    seen = {}
    orig = torch.autograd.function.FunctionCtx.save_for_backward

    def spy(ctx, *ts):
        seen["saved"] = ts
        return orig(ctx, *ts)
    torch.autograd.function.FunctionCtx.save_for_backward = spy
    try:
        y, sc = Fn.RealFormerEncoderFn.apply(x, mask, prev, 8, drop, drop, seed, *params)
    finally:
        torch.autograd.function.FunctionCtx.save_for_backward = orig
        os.environ.pop("MMVQA_RF_ENCODER", None)
    return y, sc, seen["saved"]


@pytest.mark.parametrize("B,T,L,with_prev", [(16, 28, 12, False), (3, 28, 2, True), (1, 5, 1, False), (5, 32, 3, False),
                                             (14, 17, 2, True)])
def test_cluster_kernel_matches_the_operator_chain(B, T, L, with_prev):
    """every tensor the backward pass consumes, the output, and the gradients of a full backward.  Tolerances are those of
    two bf16 evaluation orders of the same arithmetic (the FF2 reduction order differs: split-K slabs vs one
    accumulator): 3e-2 of the tensor's range on activations, 1e-4 relative on fp32 scores."""
    assert ops.rf_encoder_supported(B, T, 768, 8, 3072, L)
    with mmvqa_b200.compute_dtype_scope(torch.bfloat16):
        Fn.invalidate_weight_cache()
        blocks = _blocks(L)
        params = []
        for b in blocks:
            params.extend(block_params(b))
        x = torch.randn(B, T, 768, device="cuda").bfloat16()
        mask = torch.ones(B, T, device="cuda")
        for i in range(B):
            mask[i, T - (i % min(T, 5)):] = 0.0
        prev = torch.randn(B, 8, T, T, device="cuda") if with_prev else None
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ya, sa, sva = _run(False, xa, mask, prev, params)
        yb, sb, svb = _run(True, xb, mask, prev, params)
        for l in range(L):
            for k, nm in enumerate(NAMES):
                ta, tb = sva[l * 13 + k].float(), svb[l * 13 + k].float()
                assert torch.isfinite(tb).all(), (l, nm)
                scale = max(1.0 if nm == "scores" else 1e-3, ta.abs().max().item())
                err = (ta - tb).abs().max().item() / scale
                assert err <= (1e-4 if nm == "scores" and l == 0 else 3e-2), (l, nm, err)
        assert ((ya.float() - yb.float()).abs().max() / ya.float().abs().max()).item() < 3e-2
        assert ((sa - sb).abs().max() / sa.abs().max().clamp_min(1.0)).item() < 2e-2
        g = torch.randn(B, T, 768, device="cuda").bfloat16()
        ga, gb = [], []
        for y, xin, out in ((ya, xa, ga), (yb, xb, gb)):
            for p in params:
                p.grad = None
            y.backward(g)
            out.extend([p.grad.clone() for p in params] + [xin.grad.clone()])
        for u, v in zip(ga, gb):
            assert ((u.float() - v.float()).abs().max() / u.float().abs().max().clamp_min(1e-6)).item() < 1e-1
        Fn.invalidate_weight_cache()


def test_cluster_kernel_matches_the_oracle():
    """against the CPU restatement of the reference (oracle/mmbert_oracle.py), bf16 tolerance of the full-model tests."""
    import oracle.mmbert_oracle as O
    B, T, L = 6, 28, 4
    with mmvqa_b200.compute_dtype_scope(torch.bfloat16):
        Fn.invalidate_weight_cache()
        blocks = _blocks(L, seed=3).eval()
        x = torch.randn(B, T, 768, device="cuda")
        mask = torch.ones(B, T, device="cuda", dtype=torch.long)
        mask[2, 20:] = 0
        os.environ["MMVQA_RF_ENCODER"] = "1"
        n0 = mmvqa_b200._lib.launch_count()
        try:
            y, prev = run_blocks(list(blocks), x, None, mask, False)
        finally:
            os.environ.pop("MMVQA_RF_ENCODER", None)
        # the input cast + one bf16 cast per weight (the cache was invalidated above) + ONE encoder launch
        assert mmvqa_b200._lib.launch_count() - n0 <= 4 * L + 3, "the encoder did not take the one-launch path"
        sd = {k: v.detach().cpu().float() for k, v in blocks.state_dict().items()}
        xo, po = x.cpu(), None
        for l in range(L):
            xo, po = O.realformer_block(xo, po, mask.cpu(), sd, "%d." % l, heads=8)
        rng = xo.abs().max().item()
        assert (y.float().cpu() - xo).abs().max().item() <= 5e-2 * rng
        # RealFormer scores carry -10000 * n_layers on masked query rows: compare the unmasked rows
        keep = mask.cpu().bool()[:, :, None, None].expand_as(po)
        d = (prev.float().cpu() - po)[keep]
        assert d.abs().max().item() <= 5e-2 * po[keep].abs().max().item()
        Fn.invalidate_weight_cache()


def test_cluster_kernel_dropout_masks_match_the_backward():
    """dropout on: the forward masks are the counter hash the per-operator backward regenerates (seed + 2l / + 2l + 1,
    index row * 768 + feature): zeros of the dropped branch must be where mmvqa_dropout puts them."""
    B, T, L = 4, 28, 1
    with mmvqa_b200.compute_dtype_scope(torch.bfloat16):
        Fn.invalidate_weight_cache()
        blocks = _blocks(L, drop=0.5)
        params = list(block_params(blocks[0]))
        x = torch.randn(B, T, 768, device="cuda").bfloat16()
        _, _, sv = _run(True, x, None, None, params, drop=0.5, seed=77)
        xin, att, y1 = sv[0].float(), sv[3], sv[4].float()
        branch = y1 - xin.reshape(B * T, 768)                      # dropout(proj(att)) (up to bf16 rounding of y1)
        ones = torch.ones(B * T, 768, device="cuda").bfloat16()
        keep = ops.dropout(ones, 0.5, 77) != 0                    # seed + 2 * 0
        dropped = branch[~keep].abs()
        assert dropped.max().item() <= 2e-2 * max(1.0, xin.abs().max().item()), "a dropped element carries the branch"
        assert (branch[keep].abs() > 1e-3).float().mean().item() > 0.5
        Fn.invalidate_weight_cache()


def test_unsupported_shapes_take_the_operator_chain():
    assert not ops.rf_encoder_supported(16, 75, 768, 8, 3072, 12)     # T > 32
    assert not ops.rf_encoder_supported(64, 28, 768, 8, 3072, 12)     # too many sample groups to be co-resident
    assert not ops.rf_encoder_supported(16, 28, 128, 8, 512, 2)       # other widths


@pytest.mark.parametrize("B,T,L,with_prev,drop", [(16, 28, 3, False, 0.0), (3, 28, 2, True, 0.0), (1, 9, 1, False, 0.0),
                                                  (5, 32, 2, False, 0.0), (4, 28, 2, False, 0.3)])
def test_attention_block_backward_cluster_kernel(B, T, L, with_prev, drop):
    """csrc/rf_attn_block.cu (LN1 backward + proj dgrad + attention backward + kqv dgrad in one cluster launch) against
    the four per-operator launches it replaces: every parameter gradient, the input gradient and the gradient of the
    incoming scores.  Same bf16 arithmetic, different reduction orders: 5e-2 of each tensor's max (dropout on: the same
    counter-hash masks are regenerated on both sides)."""
    assert ops.rf_attn_block_bwd_supported(B, T, 768, 8)
    with mmvqa_b200.compute_dtype_scope(torch.bfloat16):
        Fn.invalidate_weight_cache()
        blocks = _blocks(L, seed=5, drop=drop)
        params = []
        for b in blocks:
            params.extend(block_params(b))
        x = torch.randn(B, T, 768, device="cuda").bfloat16()
        mask = torch.ones(B, T, device="cuda")
        for i in range(B):
            mask[i, T - (i % min(T, 5)):] = 0.0
        prev = torch.randn(B, 8, T, T, device="cuda") if with_prev else None
        g = torch.randn(B, T, 768, device="cuda").bfloat16()
        gs = torch.randn(B, 8, T, T, device="cuda") * 0.1
        res = {}
        for mode in ("0", "1"):
            os.environ["MMVQA_RF_ATTN_BWD"] = mode
            try:
                xin = x.clone().requires_grad_(True)
                pv = None if prev is None else prev.clone().requires_grad_(True)
                for p in params:
                    p.grad = None
                n0 = mmvqa_b200._lib.launch_count()
                y, sc, _ = _run(False, xin, mask, pv, params, drop=drop, seed=4321)
                n1 = mmvqa_b200._lib.launch_count()
                torch.autograd.backward([y, sc], [g, gs])
                n2 = mmvqa_b200._lib.launch_count()
                res[mode] = ([p.grad.clone() for p in params] + [xin.grad.clone()] + ([pv.grad.clone()] if pv is not None else []),
                             n2 - n1)
            finally:
                os.environ.pop("MMVQA_RF_ATTN_BWD", None)
        assert res["1"][1] <= res["0"][1] - 3 * L, "the cluster kernel must replace four launches per layer with one"
        for i, (u, v) in enumerate(zip(res["0"][0], res["1"][0])):
            assert torch.isfinite(v.float()).all(), i
            e = ((u.float() - v.float()).abs().max() / u.float().abs().max().clamp_min(1e-6)).item()
            assert e < 5e-2, (i, e)
        Fn.invalidate_weight_cache()
