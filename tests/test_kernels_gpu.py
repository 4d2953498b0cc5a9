"""Per-kernel parity tests (GPU): every C-ABI entry point against the CPU oracle / plain torch fp32
on the same seeded inputs.  fp32 path: tight tolerances.  bf16 path: inputs are rounded to bf16
first so that only accumulation order and output rounding differ (tolerance 2^-7 relative).
"""
import math

import pytest
import torch

from oracle import mmbert_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from mmvqa_b200 import ops
    from mmvqa_b200._lib import (ACT_GELU, ACT_NONE, ACT_RELU, ACT_SERF, EPI_ACT, EPI_ACT_ROWSUM, EPI_DACT,
                                 EPI_DACT_SCALE, EPI_RESIDUAL, EPI_STORE, MMVQAError)

DEV = "cuda"
ACTS = {"serf": 1, "gelu": 2, "relu": 3, "none": 0}
DTS = [pytest.param(torch.float32, id="fp32"), pytest.param(torch.bfloat16, id="bf16")]


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale)


def close(a, b, rtol, atol, msg=""):
    torch.testing.assert_close(a.float().cpu(), b.float().cpu(), rtol=rtol, atol=atol, msg=lambda m: f"{msg}: {m}")


def oracle_act(name, x):
    x = x.clone().requires_grad_(True)
    y = O.activation(name, x)
    (g,) = torch.autograd.grad(y.sum(), x)
    return y.detach(), g


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("act", ["serf", "gelu", "relu"])
@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("shape", [(7, 768), (5, 13)])
def test_bias_act(act, dt, shape):
    x = rnd(*shape, scale=3.0, seed=1)
    x[0, :6] = torch.tensor([-100.0, -20.0, 0.0, 49.0, 51.0, 80.0])
    bias = rnd(shape[1], seed=2)
    dy = rnd(*shape, seed=3)
    xq = x.to(dt).float()
    dyq = dy.to(dt).float()
    y_ref, g_ref = oracle_act(act, xq + bias)
    y = ops.bias_act_fwd(xq.to(dt).to(DEV), bias.to(DEV), ACTS[act])
    dx = ops.bias_act_bwd(xq.to(dt).to(DEV), bias.to(DEV), dyq.to(dt).to(DEV), ACTS[act])
    tol = dict(rtol=2e-6, atol=1e-6) if dt == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    close(y, y_ref, **tol, msg="fwd")
    close(dx, g_ref * dyq, **tol, msg="bwd")


def test_serf_known_answers_fp32():
    # SURVEY.md section 4 known answers (fp64 values, fp32 kernel)
    xs = torch.tensor([-3.0, -1.0, 0.0, 0.5, 1.0, 3.0, 10.0, 60.0])
    ys = torch.tensor([-0.164345530555, -0.342247955389, 0.0, 0.415829311482, 0.936721915472, 2.99995132254, 10.0, 60.0])
    gs = torch.tensor([-0.105382706277, 0.0671456785693, 0.673041289743, 0.967635844196, 1.08374940446, 1.00028038855,
                       1.0, 1.0])
    y = ops.bias_act_fwd(xs.to(DEV), None, ACT_SERF)
    g = ops.bias_act_bwd(xs.to(DEV), None, torch.ones_like(xs).to(DEV), ACT_SERF)
    close(y, ys, 2e-6, 1e-7)
    close(g, gs, 3e-6, 1e-7)
    assert torch.isfinite(ops.bias_act_fwd(torch.tensor([-100.0], device=DEV), None, ACT_SERF)).all()


@pytest.mark.parametrize("act", ["serf", "relu"])
def test_projector_weight_gradient_inside_the_forward_kernel(act):
    """vistok_fwd_pgrad (P = sum_hw act' f finished on chip, dW = sum_b dv P / HW) == the act'-saving path (same bf16
    rounding of act', fp32 accumulation in another order) and close to the oracle (image_encoding.py:100-115 + autograd),
    incl. a ragged last pixel tile (36 x 36 = 20 x 64 + 16) and two 128-row output tiles."""
    import mmvqa_b200
    from mmvqa_b200 import functional as Fn
    mmvqa_b200.set_compute_dtype("bf16")
    try:
        B, hidden = 3, 256
        shapes = [(24, 40), (48, 36), (24, 16), (80, 12), (512, 3)]     # the last two keep the act'-saving path
        feats = [rnd(B, c, s, s, seed=400 + i).abs().bfloat16().float().to(DEV) for i, (c, s) in enumerate(shapes)]
        ws = [(rnd(hidden, c, 1, 1, seed=410 + i) * c ** -0.5).to(DEV).requires_grad_(True) for i, (c, _) in enumerate(shapes)]
        go = rnd(len(shapes), B, hidden, seed=420).to(DEV)
        res = {}
        for flag in (False, True):
            Fn._VISTOK_PG = flag
            for w in ws:
                w.grad = None
            n0 = mmvqa_b200.launch_count()
            v = Fn.vistok_project_all(feats, ws, ACTS[act])
            v.backward(go)
            res[flag] = (v.detach().clone(), [w.grad.detach().clone() for w in ws], mmvqa_b200.launch_count() - n0)
        (v0, g0, _), (v1, g1, _) = res[False], res[True]
        close(v1, v0, 1e-5, 1e-6, msg="tokens")
        for n, (a, b) in enumerate(zip(g1, g0)):
            # (levels 0-2: both paths read SERF' from the same table -> order of summation only; level 3 (144 pixels) takes the
            # MUFU formulas on the act'-saving path, so a few act' values round to the neighbouring bf16)
            close(a, b, 1e-3 if n == 3 else 1e-4, (2e-4 if n == 3 else 2e-5) * float(b.abs().max()),
                  msg="dW level %d vs the act'-saving path" % n)
        for n, (f, w) in enumerate(zip(feats, ws)):
            wl = w.detach().reshape(hidden, -1).bfloat16().float().cpu().requires_grad_(True)
            Y = torch.einsum("mc,bcn->bmn", wl, f.flatten(2).cpu())
            vr = O.activation(act, Y).mean(-1)
            (gr,) = torch.autograd.grad(vr, wl, go[n].cpu())
            close(v1[n], vr.detach(), 1e-2, 2e-3, msg="level %d tokens vs oracle" % n)
            close(g1[n].reshape(hidden, -1), gr, 2e-2, 2e-2 * float(gr.abs().max()), msg="level %d dW vs oracle" % n)
    finally:
        Fn._VISTOK_PG = True
        mmvqa_b200.set_compute_dtype("bf16")


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("with_parts", [False, True])
def test_layernorm_bwd_deferred_matches_atomics(dt, with_parts):
    """per-CTA partial column sums + ln_partials_reduce == the atomics flush of layernorm_bwd (same dx bit for bit)."""
    rows, cols = 448, 768
    x = rnd(rows, cols, seed=300).to(dt).to(DEV)
    dy = rnd(rows, cols, seed=301).to(dt).to(DEV)
    g, b_ = (1 + 0.1 * rnd(cols, seed=302)).to(DEV), (0.1 * rnd(cols, seed=303)).to(DEV)
    _, _, mean, rstd = ops.add_layernorm_fwd(x, None, g, b_, 1e-5, False)
    parts = rnd(3, rows, cols, seed=304).to(DEV) if with_parts else None
    for p in (0.0, 0.1):
        dg0, db0, ds0 = (torch.zeros(cols, device=DEV) for _ in range(3))
        if with_parts:
            ref = ops.layernorm_bwd_parts(parts, dy, x, g, mean, rstd, dg0, db0, want_drop=p > 0, dxsum=ds0, dropout_p=p,
                                          dropout_seed=11)
        else:
            ref = ops.layernorm_bwd(dy, x, g, mean, rstd, None, dg0, db0, want_drop=p > 0, dxsum=ds0, dropout_p=p,
                                    dropout_seed=11)
        out = ops.layernorm_bwd_deferred(dy, parts, x, g, mean, rstd, want_drop=p > 0, dropout_p=p, dropout_seed=11)
        assert out is not None
        dx, dxd, partials = out
        dg1, db1, ds1 = (torch.zeros(cols, device=DEV) for _ in range(3))
        ops.ln_partials_reduce(partials, dg1, db1, ds1)
        if p > 0:
            assert torch.equal(dx, ref[0]) and torch.equal(dxd, ref[1])
        else:
            assert torch.equal(dx, ref) and dxd is None
        for a, b in ((dg1, dg0), (db1, db0), (ds1, ds0)):
            close(a, b, 1e-4, 1e-3)
    assert ops.layernorm_bwd_deferred(dy[:, :70].contiguous(), None, x[:, :70].contiguous(), g[:70], mean, rstd) is None


def test_cast_pad_multi_and_l2_prefetch():
    """one launch for the five pyramid levels (aligned, padded and odd leading dimensions) == per-tensor casts; the L2
    prefetch hint touches nothing."""
    shapes = [(16 * 24, 12544), (16 * 48, 3136), (16 * 176, 196), (16 * 512, 49), (3, 5), (7, 8), (2, 1), (4, 20), (9, 33)]
    srcs = [rnd(r, c, seed=200 + i).to(DEV) for i, (r, c) in enumerate(shapes)]
    srcs[2], srcs[3], srcs[8] = srcs[2].bfloat16(), srcs[3].bfloat16(), srcs[8].bfloat16()    # bf16 sources: pure re-pad
    lds = [(c + 7) // 8 * 8 for _, c in shapes]
    outs = ops.cast_pad_multi(srcs, lds)
    assert len(outs) == len(shapes)
    for (r, c), ld, x, y in zip(shapes, lds, srcs, outs):
        assert y.shape == (r, ld) and y.dtype == torch.bfloat16
        assert torch.equal(y[:, :c], x.to(torch.bfloat16)), (r, c)
        assert float(y[:, c:].float().abs().sum()) == 0.0
    before = [x.clone() for x in srcs[:3]]
    ops.l2_prefetch(srcs[:3] + [None], 4)
    ops.l2_prefetch(srcs + outs)            # more ranges than one list holds
    torch.cuda.synchronize()
    for a, b in zip(before, srcs[:3]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("dt", DTS)
def test_colsum_cast_dropout(dt):
    x = rnd(333, 70, seed=4).to(dt)
    s = ops.colsum(x.to(DEV), 333, 70)
    close(s, x.float().sum(0), 1e-5, 1e-4)
    a = rnd(1000, seed=5)
    close(ops.cast(a.to(DEV), torch.bfloat16), a.to(torch.bfloat16), 0, 0)
    close(ops.cast(a.to(torch.bfloat16).to(DEV), torch.float32), a.to(torch.bfloat16).float(), 0, 0)
    pad = ops.cast_pad(a.view(10, 100).to(DEV), 10, 100, 100, torch.bfloat16, 104)
    close(pad[:, :100], a.view(10, 100).to(torch.bfloat16), 0, 0)
    assert (pad[:, 100:] == 0).all()
    y = torch.ones(1 << 16, device=DEV, dtype=dt)
    d1 = ops.dropout(y, 0.3, 1234)
    d2 = ops.dropout(y, 0.3, 1234)
    assert torch.equal(d1, d2)
    keep = (d1 != 0).float().mean().item()
    assert abs(keep - 0.7) < 0.01
    close(d1[d1 != 0].float().mean(), torch.tensor(1 / 0.7), 1e-2, 0)
    z = torch.full((100,), 2.0, device=DEV, dtype=dt)
    ops.scale_(z, torch.tensor([0.5], device=DEV), 3.0)
    close(z, torch.full((100,), 3.0), 0, 0)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("cols,eps", [(768, 1e-5), (768, 1e-12), (64, 1e-5), (36, 1e-12), (1500, 1e-5)])
def test_layernorm(dt, cols, eps):
    rows = 37
    x = rnd(rows, cols, seed=6).to(dt).float()
    r = rnd(rows, cols, seed=7).to(dt).float()
    gamma = 1 + 0.1 * rnd(cols, seed=8)
    beta = 0.1 * rnd(cols, seed=9)
    dy = rnd(rows, cols, seed=10).to(dt).float()
    extra = rnd(rows, cols, seed=11).to(dt).float()
    for res in (r, None):
        xs = (x + res) if res is not None else x
        if dt == torch.bfloat16 and res is not None:
            xs = xs.to(dt).float()       # the kernel normalises the stored (rounded) sum
        xs_l = xs.clone().requires_grad_(True)
        gl, bl = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        y_ref = O.layer_norm(xs_l, gl, bl, eps)
        gx, gg, gb = torch.autograd.grad(y_ref, [xs_l, gl, bl], dy)
        y, xsum, mean, rstd = ops.add_layernorm_fwd(x.to(dt).to(DEV), None if res is None else res.to(dt).to(DEV),
                                                    gamma.to(DEV), beta.to(DEV), eps, want_sum=res is not None)
        tol = dict(rtol=1e-5, atol=2e-5) if dt == torch.float32 else dict(rtol=2e-2, atol=2e-2)
        close(y, y_ref, **tol, msg="ln fwd")
        if res is not None:
            close(xsum, xs, 0 if dt == torch.bfloat16 else 1e-6, 1e-6, msg="xsum")
        close(mean, xs.mean(-1), 1e-5, 1e-5)
        dg = torch.zeros(cols, device=DEV)
        db = torch.zeros(cols, device=DEV)
        xs_dev = xs.to(dt).to(DEV)
        dx = ops.layernorm_bwd(dy.to(dt).to(DEV), xs_dev, gamma.to(DEV), mean, rstd, extra.to(dt).to(DEV), dg, db)
        btol = dict(rtol=1e-4, atol=1e-4) if dt == torch.float32 else dict(rtol=3e-2, atol=3e-2)
        close(dx, gx + extra, **btol, msg="ln dx")
        close(dg, gg, 1e-4, 1e-3 if dt == torch.float32 else 5e-2, msg="dgamma")
        close(db, gb, 1e-4, 1e-4, msg="dbeta")
        # fused extras: dropout(dx) with the library mask and its column sums (bias gradient of the branch)
        dg2, db2, dsum = torch.zeros(cols, device=DEV), torch.zeros(cols, device=DEV), torch.zeros(cols, device=DEV)
        dx2, dxd = ops.layernorm_bwd(dy.to(dt).to(DEV), xs_dev, gamma.to(DEV), mean, rstd, extra.to(dt).to(DEV), dg2, db2,
                                     want_drop=True, dxsum=dsum, dropout_p=0.25, dropout_seed=5)
        assert torch.equal(dx2, dx)
        assert torch.equal(dxd, ops.dropout(dx, 0.25, 5))
        close(dsum, dxd.float().sum(0), 1e-4, 1e-3, msg="dxsum")
        dsum0 = torch.zeros(cols, device=DEV)
        ops.layernorm_bwd(dy.to(dt).to(DEV), xs_dev, gamma.to(DEV), mean, rstd, None, dg2, db2, dxsum=dsum0)
        dx_plain = ops.layernorm_bwd(dy.to(dt).to(DEV), xs_dev, gamma.to(DEV), mean, rstd, None, dg2, db2)
        close(dsum0, dx_plain.float().sum(0), 1e-4, 1e-3, msg="dxsum (no dropout)")


# ------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------
def _gemm_case(dt, M, N, K, a_trans, b_trans, seed=0):
    A = rnd(M, K, seed=seed).to(dt)
    B = rnd(N, K, seed=seed + 1).to(dt)
    ref = A.float() @ B.float().t()
    As = A.t().contiguous() if a_trans else A
    Bs = B.t().contiguous() if b_trans else B
    return As.to(DEV), (M if a_trans else K), Bs.to(DEV), (N if b_trans else K), ref


GEMM_SHAPES = [(448, 768, 768), (448, 3072, 768), (130, 200, 96), (3584, 288, 96), (64, 64, 64), (1, 8, 8),
               (300, 1552, 768), (257, 72, 1000)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("a_trans,b_trans", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("dt", DTS)
def test_gemm_store(dt, M, N, K, a_trans, b_trans):
    if dt == torch.bfloat16 and ((a_trans and M % 8) or (b_trans and N % 8) or K % 8):
        pytest.skip("TMA needs 16-byte row strides")
    A, lda, B, ldb, ref = _gemm_case(dt, M, N, K, a_trans, b_trans)
    bias = rnd(N, seed=5).to(DEV)
    for cdt in ([torch.float32] if dt == torch.float32 else [torch.bfloat16, torch.float32]):
        C = torch.full((M, N), float("nan"), device=DEV, dtype=cdt)
        ops.gemm(M, N, K, A, lda, a_trans, B, ldb, b_trans, C, N, bias=bias)
        tol = dict(rtol=1e-4, atol=1e-3) if dt == torch.float32 else (
            dict(rtol=1e-2, atol=1e-1) if cdt == torch.bfloat16 else dict(rtol=1e-3, atol=2e-2))
        close(C, ref + bias.cpu(), **tol, msg=f"gemm {dt} {cdt} {a_trans}{b_trans}")


@pytest.mark.parametrize("dt", DTS)
def test_gemm_bf16_inputs_exact_small_ints(dt):
    # integer-valued operands: every product and partial sum is exact in fp32 -> results must be bit-exact,
    # which pins the swizzle / descriptor / k-advance arithmetic of the tensor-core path
    g = torch.Generator().manual_seed(3)
    for (M, N, K, at, bt) in [(256, 256, 256, False, False), (128, 64, 64, False, True), (128, 128, 192, True, True),
                              (384, 320, 128, True, False), (200, 136, 72, False, True)]:
        A = torch.randint(-3, 4, (M, K), generator=g).float()
        B = torch.randint(-3, 4, (N, K), generator=g).float()
        ref = A @ B.t()
        As = (A.t().contiguous() if at else A).to(dt).to(DEV)
        Bs = (B.t().contiguous() if bt else B).to(dt).to(DEV)
        C = torch.empty(M, N, device=DEV, dtype=torch.float32)
        ops.gemm(M, N, K, As, M if at else K, at, Bs, N if bt else K, bt, C, N)
        assert torch.equal(C.cpu(), ref), f"{dt} {(M, N, K, at, bt)} max err {(C.cpu() - ref).abs().max()}"


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("act", ["serf", "gelu", "relu"])
def test_gemm_epilogues(dt, act):
    M, N, K = 200, 264, 136
    A, lda, B, ldb, ref = _gemm_case(dt, M, N, K, False, False, seed=20)
    ref = ref * 0.2
    A = (A.float() * 0.2).to(dt)
    ref = A.float().cpu() @ B.float().cpu().t()
    bias = rnd(N, seed=21)
    pre_ref = ref + bias
    tol = dict(rtol=1e-4, atol=1e-3) if dt == torch.float32 else dict(rtol=2e-2, atol=5e-2)
    # EPI_ACT with pre-activation side output
    C = torch.empty(M, N, device=DEV, dtype=dt)
    pre = torch.empty(M, N, device=DEV, dtype=dt)
    ops.gemm(M, N, K, A, lda, False, B, ldb, False, C, N, bias=bias.to(DEV), epilogue=EPI_ACT, act=ACTS[act], aux_out=pre,
             ld_aux_out=N)
    close(pre, pre_ref, **tol, msg="preact")
    close(C, O.activation(act, pre_ref), **tol, msg="act")
    # EPI_RESIDUAL
    res = rnd(M, N, seed=22).to(dt)
    ops.gemm(M, N, K, A, lda, False, B, ldb, False, C, N, bias=bias.to(DEV), epilogue=EPI_RESIDUAL, aux_in=res.to(DEV),
             ld_aux_in=N)
    close(C, pre_ref + res.float(), **tol, msg="residual")
    # EPI_RESIDUAL + dropout: identical mask to mmvqa_dropout
    ops.gemm(M, N, K, A, lda, False, B, ldb, False, C, N, bias=bias.to(DEV), epilogue=EPI_RESIDUAL, aux_in=res.to(DEV),
             ld_aux_in=N, dropout_p=0.25, dropout_seed=77)
    keep = ops.dropout(torch.ones(M, N, device=DEV), 0.25, 77).cpu()
    close(C, pre_ref * keep + res.float(), **tol, msg="residual+dropout")
    # EPI_DACT
    aux = rnd(M, N, seed=23, scale=2.0).to(dt)
    _, dact = oracle_act(act, aux.float())
    cs = torch.zeros(N, device=DEV)
    ops.gemm(M, N, K, A, lda, False, B, ldb, False, C, N, epilogue=EPI_DACT, act=ACTS[act], aux_in=aux.to(DEV), ld_aux_in=N,
             colsum_out=cs)
    close(C, ref * dact, **tol, msg="dact")
    close(cs, (ref * dact).sum(0), 1e-3 if dt == torch.float32 else 2e-2, 1e-2 if dt == torch.float32 else 0.3, msg="fused colsum")


@pytest.mark.parametrize("dt", DTS)
def test_gemm_splitk_accumulate_and_batch(dt):
    M, N, K = 136, 72, 1000
    A, lda, B, ldb, ref = _gemm_case(dt, M, N, K, True, True, seed=30)
    C = torch.zeros(M, N, device=DEV)
    ops.gemm(M, N, K, A, lda, True, B, ldb, True, C, N, accumulate=True, split_k=5)
    close(C, ref, 1e-3, 2e-2, msg="split-k")
    # batched sum over batch: C = sum_b A_b B_b^T
    Bn, M, N, K = 3, 96, 40, 200
    A = rnd(Bn, M, K, seed=31).to(dt)
    B = rnd(Bn, N, K, seed=32).to(dt)
    ref = torch.einsum("bmk,bnk->mn", A.float(), B.float())
    C = torch.zeros(M, N, device=DEV)
    ops.gemm(M, N, K, A.to(DEV), K, False, B.to(DEV), K, False, C, N, accumulate=True, batch=Bn, a_batch_rows=M, b_batch_rows=N)
    close(C, ref, 1e-3, 2e-2, msg="batched accumulate")
    # batched store with shared A
    W = rnd(M, K, seed=33).to(dt)
    ref = torch.einsum("mk,bnk->bmn", W.float(), B.float())
    C = torch.empty(Bn, M, N, device=DEV, dtype=dt)
    ops.gemm(M, N, K, W.to(DEV), K, False, B.to(DEV), K, False, C, N, batch=Bn, a_batch_rows=0, b_batch_rows=N,
             c_batch_stride=M * N)
    close(C, ref, 1e-2 if dt == torch.bfloat16 else 1e-4, 1e-1 if dt == torch.bfloat16 else 1e-3, msg="batched store")


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("act", ["serf", "relu"])
def test_gemm_projector_epilogues(dt, act):
    # visual-token pooling: v[b, m] = mean_n act(W f_b)[m, n]  with f stored [C, HW] (MN-major B operand)
    Bn, hidden, Cc, HW = 3, 136, 24, 49
    ld = 56 if dt == torch.bfloat16 else HW
    W = (rnd(hidden, Cc, seed=40) * 0.3).to(dt)
    f = rnd(Bn, Cc, HW, seed=41).abs().to(dt)
    fpad = torch.zeros(Bn, Cc, ld, dtype=dt)
    fpad[:, :, :HW] = f
    Y = torch.einsum("mc,bcn->bmn", W.float(), f.float())
    v_ref = O.activation(act, Y).mean(-1)
    v = torch.zeros(Bn, hidden, device=DEV)
    ops.gemm(hidden, HW, Cc, W.to(DEV), Cc, False, fpad.to(DEV), ld, True, None, 0, epilogue=EPI_ACT_ROWSUM, act=ACTS[act],
             rowsum_out=v, scale=1.0 / HW, batch=Bn, a_batch_rows=0, b_batch_rows=Cc)
    close(v, v_ref, 1e-4 if dt == torch.float32 else 1e-2, 1e-5 if dt == torch.float32 else 1e-2, msg="rowsum")
    dv = rnd(Bn, hidden, seed=42)
    _, dact = oracle_act(act, Y)
    G_ref = dact * dv[:, :, None] / HW
    G = torch.zeros(Bn, hidden, ld, device=DEV, dtype=dt)
    ops.gemm(hidden, HW, Cc, W.to(DEV), Cc, False, fpad.to(DEV), ld, True, G, ld, epilogue=EPI_DACT_SCALE, act=ACTS[act],
             rowscale=dv.to(DEV), scale=1.0 / HW, batch=Bn, a_batch_rows=0, b_batch_rows=Cc, c_batch_stride=hidden * ld)
    close(G[:, :, :HW], G_ref, 1e-4 if dt == torch.float32 else 2e-2, 1e-6 if dt == torch.float32 else 2e-3, msg="dact_scale")


def test_gemm_argument_errors():
    A = torch.zeros(8, 8, device=DEV)
    with pytest.raises(MMVQAError):
        ops.gemm(8, 8, 8, A, 4, False, A, 8, False, A, 8)            # lda too small
    with pytest.raises(MMVQAError):
        ops.gemm(8, 8, 8, A, 8, False, A, 8, False, A, 8, epilogue=EPI_RESIDUAL)   # missing aux_in
    Ab = torch.zeros(8, 12, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(MMVQAError):
        ops.gemm(8, 8, 12, Ab, 12, False, Ab, 12, False, torch.zeros(8, 8, device=DEV), 8)   # 24-byte rows: TMA refuses
    with pytest.raises(MMVQAError):
        ops.bias_act_fwd(torch.zeros(4), None, ACT_SERF)             # CPU tensor: no CPU fallback


# ------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------
def _mask(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    m = torch.ones(B, T)
    for b in range(B):
        n = int(torch.randint(max(1, T // 3), T + 1, (1,), generator=g))
        m[b, n:] = 0
    return m


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,T,heads,d", [(3, 28, 8, 96), (2, 75, 8, 96), (2, 128, 2, 64), (2, 10, 8, 8), (1, 1, 1, 4)])
def test_rf_attention(dt, B, T, heads, d):
    H = heads * d
    kqv = (rnd(B, T, heads, 3 * d, seed=50) * 0.5).to(dt)
    prev = rnd(B, T, T, heads, seed=51) * 0.3
    mask = _mask(B, T, 52)
    dout = rnd(B, T, H, seed=53).to(dt)
    dsc = rnd(B, T, T, heads, seed=54) * 0.1
    kq = kqv.float().clone().requires_grad_(True)
    pv = prev.clone().requires_grad_(True)
    k, q, v = kq[..., :d], kq[..., d:2 * d], kq[..., 2 * d:]
    s = torch.einsum("bihk,bjhk->bijh", q, k) / math.sqrt(d) + pv
    s = s - 10000.0 * (1.0 - mask)[:, :, None, None]
    att = torch.softmax(s, dim=2)
    o = torch.einsum("btih,bihs->bths", att, v).reshape(B, T, H)
    gk, gp = torch.autograd.grad([o, s], [kq, pv], [dout.float(), dsc])
    prev_n = prev.permute(0, 3, 1, 2).contiguous().to(DEV)
    out, scores = ops.rf_attn_fwd(kqv.reshape(B * T * heads, 3 * d).to(DEV), prev_n, mask.to(DEV), B, T, heads, d)
    tol = dict(rtol=1e-4, atol=1e-4) if dt == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    close(scores.permute(0, 2, 3, 1), s.detach(), 1e-5, 1e-3, msg="scores")
    close(out.view(B, T, H), o.detach(), **tol, msg="out")
    dkqv, dprev = ops.rf_attn_bwd(kqv.reshape(B * T * heads, 3 * d).to(DEV), scores, dout.reshape(B * T, H).to(DEV),
                                  dsc.permute(0, 3, 1, 2).contiguous().to(DEV), True, B, T, heads, d)
    close(dprev.permute(0, 2, 3, 1), gp, 1e-3 if dt == torch.float32 else 3e-2, 1e-4 if dt == torch.float32 else 2e-2, msg="dprev")
    close(dkqv.view(B, T, heads, 3 * d), gk, 1e-3 if dt == torch.float32 else 3e-2, 1e-4 if dt == torch.float32 else 3e-2,
          msg="dkqv")
    # no prev, no mask, no incoming score gradient
    out2, sc2 = ops.rf_attn_fwd(kqv.reshape(B * T * heads, 3 * d).to(DEV), None, None, B, T, heads, d)
    s2 = torch.einsum("bihk,bjhk->bijh", q, k).detach() / math.sqrt(d)
    close(sc2.permute(0, 2, 3, 1), s2, 1e-5, 1e-4, msg="scores (no prev)")


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,T,heads,d", [(3, 28, 12, 64), (2, 75, 12, 64), (2, 9, 4, 16)])
def test_mhsa_attention(dt, B, T, heads, d):
    H = heads * d
    qkv = (rnd(B, T, 3 * H, seed=60) * 0.5).to(dt)
    mask = _mask(B, T, 61)
    dout = rnd(B, T, H, seed=62).to(dt)
    z = qkv.float().clone().requires_grad_(True)
    q, k, v = (z[..., i * H:(i + 1) * H].view(B, T, heads, d).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-2, -1) / math.sqrt(d) - 10000.0 * (1.0 - mask[:, None, None, :])
    pr = torch.softmax(s, dim=-1)
    o = (pr @ v).transpose(1, 2).reshape(B, T, H)
    (gz,) = torch.autograd.grad(o, z, dout.float())
    out, probs = ops.mhsa_fwd(qkv.reshape(B * T, 3 * H).to(DEV), mask.to(DEV), B, T, heads, d, 0.0, 0)
    tol = dict(rtol=1e-4, atol=1e-4) if dt == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    close(probs, pr.detach(), 1e-4 if dt == torch.float32 else 2e-2, 1e-5 if dt == torch.float32 else 1e-2, msg="probs")
    close(out.view(B, T, H), o.detach(), **tol, msg="out")
    dqkv = ops.mhsa_bwd(qkv.reshape(B * T, 3 * H).to(DEV), probs, dout.reshape(B * T, H).to(DEV), B, T, heads, d, 0.0, 0)
    close(dqkv.view(B, T, 3 * H), gz, 1e-3 if dt == torch.float32 else 4e-2, 1e-4 if dt == torch.float32 else 3e-2, msg="dqkv")
    # dropout on the probabilities: statistical check + fwd/bwd mask agreement (gradient of sum(out) w.r.t. v)
    out_d, _ = ops.mhsa_fwd(qkv.reshape(B * T, 3 * H).to(DEV), mask.to(DEV), B, T, heads, d, 0.5, 9)
    assert torch.isfinite(out_d.float()).all()


# ------------------------------------------------------------------------------------------
# fusion / pooling / losses
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("shape", [(3, 12, 72, 50, 5), (16, 28, 768, 300, 5), (5, 9, 256, 40, 2)])
def test_embed_ln_scatter(dt, shape):
    # (H = 768 / 256 with bf16 take the vector backward kernel: register-accumulated token-type / gamma / beta sums)
    B, T, H, V, nvis = shape
    ids = torch.randint(0, V, (B, T), generator=torch.Generator().manual_seed(70))
    ids[:, 1:6] = 0
    seg = torch.randint(0, 2, (B, T), generator=torch.Generator().manual_seed(71))
    word, pos, typ = rnd(V, H, seed=72), rnd(T + 8, H, seed=73), rnd(2, H, seed=74)
    gamma, beta = 1 + 0.1 * rnd(H, seed=75), 0.1 * rnd(H, seed=76)
    vis = rnd(nvis, B, H, seed=77)
    dh = rnd(B, T, H, seed=78).to(dt).float()
    leaves = [t.clone().requires_grad_(True) for t in (word, pos, typ, gamma, beta, vis)]
    w_, p_, t_, g_, b_, v_ = leaves
    e = O.bert_embeddings(ids, seg, w_, p_, t_, g_, b_, 1e-12)
    h_ref = O.fuse_visual_tokens(e, [v_[n] for n in range(nvis)])
    grads = torch.autograd.grad(h_ref, leaves, dh)
    grads[0][0].zero_()                               # padding_idx = 0 receives no gradient
    h, mean, rstd = ops.embed_ln_scatter_fwd(ids.to(DEV), seg.to(DEV), word.to(DEV), pos.to(DEV), typ.to(DEV), gamma.to(DEV),
                                             beta.to(DEV), vis.to(DEV), dt, 1e-12, 0.0, 0)
    tol = dict(rtol=1e-5, atol=1e-5) if dt == torch.float32 else dict(rtol=1e-2, atol=2e-2)
    close(h, h_ref.detach(), **tol, msg="embed fwd")
    outs = [torch.zeros_like(t, device=DEV) for t in (word, pos, typ, gamma, beta)]
    dvis = torch.empty(nvis, B, H, device=DEV)
    ops.embed_ln_scatter_bwd(dh.to(dt).to(DEV), ids.to(DEV), seg.to(DEV), word.to(DEV), pos.to(DEV), typ.to(DEV),
                             gamma.to(DEV), mean, rstd, *outs, dvis, nvis, 0, 0.0, 0)
    for name, got, want in zip(["dword", "dpos", "dtyp", "dgamma", "dbeta", "dvis"], outs + [dvis], grads):
        close(got, want, 1e-4, 1e-4, msg=name)


@pytest.mark.parametrize("dt", DTS)
def test_masked_mean_and_l2norm(dt):
    B, T, H = 4, 28, 768
    h = rnd(B, T, H, seed=80).to(dt)
    mask = _mask(B, T, 81)
    mask[3] = 0                                        # fully masked sample -> clamp(1e-9) branch
    hl = h.float().clone().requires_grad_(True)
    ref = O.mean_pooling(hl, mask)
    dout = rnd(B, H, seed=82).to(dt)
    (gh,) = torch.autograd.grad(ref, hl, dout.float())
    out = ops.masked_mean_fwd(h.to(DEV), mask.to(DEV))
    tol = dict(rtol=1e-5, atol=1e-5) if dt == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    close(out, ref.detach(), **tol)
    close(ops.masked_mean_bwd(dout.to(DEV), mask.to(DEV), T), gh, **tol)
    if dt == torch.float32:
        x = rnd(9, 128, seed=83)
        xl = x.clone().requires_grad_(True)
        y_ref = xl / torch.clamp(xl.norm(dim=1, keepdim=True), min=1e-12)
        dy = rnd(9, 128, seed=84)
        (gx,) = torch.autograd.grad(y_ref, xl, dy)
        y, inv = ops.l2norm_fwd(x.to(DEV))
        close(y, y_ref.detach(), 1e-5, 1e-6)
        close(ops.l2norm_bwd(y, inv, dy.to(DEV)), gx, 1e-4, 1e-5)


def test_asl_against_golden(golden):
    g = golden("losses")
    for key in ("asl_kat", "asl_default", "asl_g1_2_eps0", "asl_sum"):
        c = g[key]
        kw = c.get("kw", {})
        lg = c["logits"].float().to(DEV)
        Bn, Cn = lg.shape
        rows, dl, tc = ops.asl_fwd_bwd(lg, Cn, c["target"].to(DEV), Cn, float(kw.get("gamma_pos", 0)),
                                       float(kw.get("gamma_neg", 4)), float(kw.get("eps", 0.1)), True, True)
        red = kw.get("reduction", "mean")
        loss = rows.mean() if red == "mean" else rows
        close(loss, c["loss"].float(), 1e-5, 1e-6, msg=key)
        scale = 1.0 / Bn if red == "mean" else 1.0
        close(dl * scale, c["grad"].float(), 1e-4, 1e-6, msg=key + " grad")
        if "targets_classes" in c:
            close(tc, c["targets_classes"].float(), 1e-6, 1e-7)


def test_ce_rows():
    rows, Cn, ld = 37, 30522, 30528
    lg = rnd(rows, Cn, seed=90, scale=2.0)
    tgt = torch.randint(0, Cn, (rows,), generator=torch.Generator().manual_seed(91))
    tgt[:5] = 0
    ll = lg.clone().requires_grad_(True)
    ref = -torch.log_softmax(ll, -1).gather(-1, tgt[:, None]).squeeze(1)
    (gref,) = torch.autograd.grad(ref.sum(), ll)
    buf = torch.zeros(rows, ld, device=DEV)
    buf[:, :Cn] = lg.to(DEV)
    dl = torch.empty(rows, ld, device=DEV)
    loss_rows = ops.ce_fwd_bwd(buf, ld, tgt.to(DEV), rows, Cn, 1.0, dl, ld)
    close(loss_rows, ref.detach(), 1e-5, 1e-5)
    close(dl[:, :Cn], gref, 1e-4, 1e-7)
    assert abs(O.mlm_nll(lg[None], tgt[None]).item() - loss_rows.mean().item()) < 1e-4


def test_supcon_rows_against_oracle(golden):
    r = golden("losses")["supcon_rand"]
    f = r["features"].float()
    bsz, nv, D = f.shape
    Fm = torch.cat([f[:, v] for v in range(nv)], 0)
    raw = (Fm @ Fm.t()).to(DEV)
    for tag, mask in (("simclr", None), ("soft", r["soft"].float())):
        rows, G = ops.supcon_rows(raw, None if mask is None else mask.to(DEV), bsz, 0, 0.07, 0.07, True)
        close(rows.mean(), r[tag + "_loss"].float(), 1e-5, 1e-5, msg=tag)
        Gs = (G / rows.numel()).cpu()
        dF = Gs @ Fm + Gs.t() @ Fm
        want = r[tag + "_grad"].float()
        want_vm = torch.cat([want[:, v] for v in range(nv)], 0)
        close(dF, want_vm, 1e-3, 1e-5, msg=tag + " grad")


@pytest.mark.parametrize("Cc,HW,hidden,Bn", [(24, 1000, 256, 3), (48, 3136, 768, 2), (80, 784, 136, 2), (128, 520, 128, 1)])
@pytest.mark.parametrize("act", ["serf", "relu"])
def test_persistent_projector_kernel(Cc, HW, hidden, Bn, act):
    """vistok.cu (bf16, C <= 128, >= 2 pixel tiles): pooled forward and the recompute backward vs torch."""
    dt = torch.bfloat16
    ld = (HW + 7) // 8 * 8
    W = (rnd(hidden, Cc, seed=100) * 0.3).to(dt)
    f = rnd(Bn, Cc, HW, seed=101).abs().to(dt)
    fpad = torch.zeros(Bn, Cc, ld, dtype=dt)
    fpad[:, :, :HW] = f
    Y = torch.einsum("mc,bcn->bmn", W.float(), f.float())
    v_ref = O.activation(act, Y).mean(-1)
    v = torch.zeros(Bn, hidden, device=DEV)
    ops.gemm(hidden, HW, Cc, W.to(DEV), Cc, False, fpad.to(DEV), ld, True, None, 0, epilogue=EPI_ACT_ROWSUM, act=ACTS[act],
             rowsum_out=v, scale=1.0 / HW, batch=Bn, a_batch_rows=0, b_batch_rows=Cc)
    close(v, v_ref, 1e-2, 2e-3, msg="pooled forward")
    dv = rnd(Bn, hidden, seed=102)
    _, dact = oracle_act(act, Y)
    # forward that also keeps act'(.) for the backward pass, then dW = sum_b (dv_b/HW) . act'_b f_b^T
    v2 = torch.zeros(Bn, hidden, device=DEV)
    actp = torch.full((Bn, hidden, ld), float("nan"), device=DEV, dtype=dt)
    ops.gemm(hidden, HW, Cc, W.to(DEV), Cc, False, fpad.to(DEV), ld, True, None, 0, epilogue=EPI_ACT_ROWSUM, act=ACTS[act],
             rowsum_out=v2, scale=1.0 / HW, batch=Bn, a_batch_rows=0, b_batch_rows=Cc, aux_out=actp, ld_aux_out=ld)
    close(v2, v_ref, 1e-2, 2e-3, msg="pooled forward (+act')")
    ap = actp[:, :, :HW].float().cpu()
    assert torch.isfinite(ap).all()
    sel = (Y.abs() > 0.05) if act == "relu" else torch.ones_like(Y, dtype=torch.bool)
    close(ap[sel], dact[sel], 2e-2, 1e-2, msg="stored act'")
    dw = torch.zeros(hidden, Cc, device=DEV)
    ops.gemm(hidden, Cc, HW, actp, ld, False, fpad.to(DEV), ld, False, dw, Cc, accumulate=True, split_k=2, batch=Bn,
             a_batch_rows=hidden, b_batch_rows=Cc, c_batch_stride=0, rowscale=dv.to(DEV), scale=1.0 / HW)
    dw_ref = torch.einsum("bmn,bcn,bm->mc", ap, f.float(), dv) / HW
    close(dw, dw_ref, 2e-2, 2e-3 * float(dw_ref.abs().max()), msg="row-scaled batched accumulate")
    G_ref = dact * dv[:, :, None] / HW
    G = torch.full((Bn, hidden, ld), float("nan"), device=DEV, dtype=dt)
    ops.gemm(hidden, HW, Cc, W.to(DEV), Cc, False, fpad.to(DEV), ld, True, G, ld, epilogue=EPI_DACT_SCALE, act=ACTS[act],
             rowscale=dv.to(DEV), scale=1.0 / HW, batch=Bn, a_batch_rows=0, b_batch_rows=Cc, c_batch_stride=hidden * ld)
    got = G[:, :, :HW].float().cpu()
    assert torch.isfinite(got).all()
    if act == "relu":       # act' flips where the bf16 product crosses 0: compare away from the kink
        sel = Y.abs() > 0.05
        close(got[sel], G_ref[sel], 2e-2, 1e-6, msg="recompute backward (relu)")
    else:
        close(got, G_ref, 2e-2, 2e-3 * float(G_ref.abs().max()), msg="recompute backward")


def test_projector_takes_bf16_feature_maps_in_place():
    """SURVEY.md section 8 row f4 (backbone hand-off): a backbone running under bf16 autocast hands the projector bf16
    NCHW maps; they are the TMA operand as they are (no cast launch).  Same tokens and the same weight gradients as the
    fp32 maps holding the same (bf16-representable) values, and close to the oracle (image_encoding.py:100-115)."""
    import mmvqa_b200
    from mmvqa_b200 import functional as Fn
    mmvqa_b200.set_compute_dtype("bf16")
    try:
        B, hidden = 2, 256
        shapes = [(24, 40), (48, 20), (80, 12), (176, 6), (512, 3)]       # odd pixel counts too: 36 and 9 need the padded ld
        feats32 = [rnd(B, c, s, s, seed=300 + i).abs().bfloat16().float().to(DEV) for i, (c, s) in enumerate(shapes)]
        ws = [(rnd(hidden, c, 1, 1, seed=310 + i) * c ** -0.5).to(DEV).requires_grad_(True) for i, (c, _) in enumerate(shapes)]
        go = rnd(len(shapes), B, hidden, seed=320).to(DEV)
        outs, grads = [], []
        Fn.vistok_project_all(feats32, ws, ACTS["serf"])          # warm the bf16 weight-operand cache: count map casts only
        for feats in (feats32, [f.bfloat16() for f in feats32]):
            for w in ws:
                w.grad = None
            n0 = mmvqa_b200.launch_count()
            v = Fn.vistok_project_all(feats, ws, ACTS["serf"])
            launches = mmvqa_b200.launch_count() - n0
            v.backward(go)
            outs.append((v.detach().clone(), launches))
            grads.append([w.grad.detach().clone() for w in ws])
        (v32, l32), (v16, l16) = outs
        # same operand bytes; the pooled sums are combined with one fp32 atomic per (row, pixel split): order only
        close(v16, v32, 1e-5, 1e-6, msg="tokens from bf16 maps vs fp32 maps")
        for g32, g16 in zip(*grads):
            close(g16, g32, 1e-5, 1e-6 * float(g32.abs().max()), msg="dW")     # split-K atomics: order only
        # fp32 maps: one multi-tensor cast launch; bf16 maps: the levels with 16-byte rows are used in place and only the
        # two odd ones (36 and 9 pixels) are re-padded, in one launch -- never more launches than the fp32 hand-off
        assert l16 <= l32, "bf16 maps must not need more launches than fp32 maps (%d vs %d)" % (l16, l32)
        n0 = mmvqa_b200.launch_count()
        Fn.vistok_project_all([f.bfloat16() for f in feats32[:3]], ws[:3], ACTS["serf"])      # 1600 / 400 / 144 pixels
        assert mmvqa_b200.launch_count() - n0 == 3, "aligned bf16 maps are the TMA operand as they are: no cast launch"
        for n, (f, w) in enumerate(zip(feats32, ws)):
            Y = torch.einsum("mc,bcn->bmn", w.detach().reshape(hidden, -1).bfloat16().float().cpu(), f.flatten(2).cpu())
            close(v16[n], O.activation("serf", Y).mean(-1), 1e-2, 2e-3, msg="level %d vs oracle" % n)
    finally:
        mmvqa_b200.set_compute_dtype("bf16")


# ------------------------------------------------------------------------------------------
# split-K slabs: GEMM partial tiles reduced by the LayerNorm pass that follows (realformer.py:49-50 at M = 448)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("M,N,K,ns,b_trans", [(448, 768, 3072, 3, False), (448, 768, 3072, 4, True), (37, 128, 1000, 5, False),
                                              (130, 72, 64, 3, False)])
def test_gemm_splitk_slabs(dt, M, N, K, ns, b_trans):
    A, lda, B, ldb, ref = _gemm_case(dt, M, N, K, False, b_trans, seed=40)
    bias = rnd(N, seed=41)
    parts = torch.full((ns, M, N), float("nan"), device=DEV)         # every slab element must be written (no zero-fill)
    ops.gemm(M, N, K, A, lda, False, B, ldb, b_trans, parts, N, bias=bias.to(DEV), split_k=ns, c_split_stride=M * N)
    assert torch.isfinite(parts).all()
    close(parts.sum(0), ref + bias, 1e-3, 3e-2 if dt == torch.bfloat16 else 2e-3, msg="sum of slabs")
    with pytest.raises(MMVQAError):                                  # split-K without accumulate needs a slab stride
        ops.gemm(M, N, K, A, lda, False, B, ldb, b_trans, parts, N, split_k=ns)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("cols,p", [(768, 0.0), (768, 0.1), (128, 0.3), (256, 0.0)])
def test_layernorm_over_splitk_partials(dt, cols, p):
    rows, ns, eps, seed = 37, 3, 1e-5, 11
    parts = torch.stack([rnd(rows, cols, seed=50 + i) for i in range(ns)]).to(DEV)
    res = rnd(rows, cols, seed=60).to(dt).to(DEV)
    gamma, beta = (1 + 0.1 * rnd(cols, seed=61)).to(DEV), (0.1 * rnd(cols, seed=62)).to(DEV)
    # reference composition with the library's own pieces: sum -> (round) -> dropout -> + residual -> LN
    tot = parts.sum(0)
    dropped = ops.dropout(tot.contiguous(), p, seed) if p > 0 else tot          # fp32 dropout of the fp32 sum
    s_ref = (dropped + res.float()).to(dt)
    y_ref, _, mean_ref, rstd_ref = ops.add_layernorm_fwd(s_ref, None, gamma, beta, eps, want_sum=False)
    y, s, mean, rstd = ops.add_layernorm_fwd_parts(parts, res, gamma, beta, eps, dt, p, seed)
    tol = dict(rtol=1e-5, atol=1e-5) if dt == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    close(s, s_ref, **tol, msg="stored sum")
    close(y, y_ref, **(dict(rtol=1e-4, atol=1e-4) if dt == torch.float32 else dict(rtol=2e-2, atol=3e-2)), msg="ln(parts)")
    close(mean, mean_ref, 1e-3, 1e-3)
    # backward: dy = sum(parts) + dy_res
    dres = rnd(rows, cols, seed=63).to(dt).to(DEV)
    dy_ref = (tot + dres.float()).to(dt)
    dg0, db0, ds0 = (torch.zeros(cols, device=DEV) for _ in range(3))
    dg1, db1, ds1 = (torch.zeros(cols, device=DEV) for _ in range(3))
    dx_ref, dxd_ref = ops.layernorm_bwd(dy_ref, s_ref, gamma, mean_ref, rstd_ref, None, dg0, db0, want_drop=True, dxsum=ds0,
                                        dropout_p=0.2, dropout_seed=3)
    dx, dxd = ops.layernorm_bwd_parts(parts, dres, s_ref, gamma, mean_ref, rstd_ref, dg1, db1, want_drop=True, dxsum=ds1,
                                      dropout_p=0.2, dropout_seed=3)
    btol = dict(rtol=1e-4, atol=1e-4) if dt == torch.float32 else dict(rtol=2e-2, atol=3e-2)
    close(dx, dx_ref, **btol, msg="dx")
    close(dxd, dxd_ref, **btol, msg="dropout(dx)")
    close(dg1, dg0, 1e-3, 5e-2, msg="dgamma")
    close(db1, db0, 1e-3, 5e-2, msg="dbeta")
    close(ds1, ds0, 1e-3, 5e-2, msg="dxsum")
    dx_nores = ops.layernorm_bwd_parts(parts, None, s_ref, gamma, mean_ref, rstd_ref, dg1, db1)
    dx_nores_ref = ops.layernorm_bwd(tot.to(dt), s_ref, gamma, mean_ref, rstd_ref, None, dg0, db0)
    close(dx_nores, dx_nores_ref, **btol, msg="dx (no residual)")


# ------------------------------------------------------------------------------------------
# caption-similarity mask (supcon_utils.py:110-138)
# ------------------------------------------------------------------------------------------
def test_jaccard_mask_against_golden_and_oracle(golden):
    import random

    from mmvqa_b200.similarity import build_mask, jaccard_mask
    for name, case in golden("jaccard").items():
        got = jaccard_mask(case["caption"], case["aug"])
        assert got.dtype == torch.float32 and torch.equal(got.cpu(), case["mask"]), name      # integer work: bit-exact
    rng = random.Random(1)
    words = [f"w{i}" for i in range(60)]
    bsz = 257                                                                                # ragged sizes, > one CTA row group
    cap = [" ".join(rng.choice(words) for _ in range(rng.randint(0, 40))) for _ in range(bsz)]
    aug = [" ".join(rng.choice(words) for _ in range(rng.randint(0, 40))) for _ in range(bsz)]
    assert torch.equal(jaccard_mask(cap, aug).cpu(), O.jaccard_mask(cap, aug, bsz))
    assert build_mask(bsz, cap, aug, "simclr") is None
    assert torch.equal(build_mask(bsz, cap, aug, "supcon").cpu(), O.jaccard_mask(cap, aug, bsz))
    with pytest.raises(MMVQAError):
        jaccard_mask(cap, aug, device="cpu")


# ------------------------------------------------------------------------------------------
# RealFormer attention with the kqv projection inside the kernel == kqv GEMM + attention kernel
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T,heads,d", [(3, 28, 8, 96), (2, 75, 8, 96), (2, 128, 2, 64), (2, 10, 8, 16)])
def test_rf_attention_fused_kqv(B, T, heads, d):
    bf = torch.bfloat16
    H = heads * d
    x = rnd(B * T, H, seed=70).to(bf).to(DEV)
    w = (rnd(3 * d, d, seed=71) / math.sqrt(d)).to(bf).to(DEV)
    prev = rnd(B, heads, T, T, seed=72).to(DEV)
    mask = torch.ones(B, T, device=DEV)
    mask[0, T - T // 3:] = 0
    kqv_ref = torch.empty(B * T * heads, 3 * d, device=DEV, dtype=bf)
    ops.gemm(B * T * heads, 3 * d, d, x, d, False, w, d, False, kqv_ref, 3 * d)
    for pv in (prev, None):
        out_ref, sc_ref = ops.rf_attn_fwd(kqv_ref, pv, mask, B, T, heads, d)
        out, sc, kqv = ops.rf_attn_fwd_fused(x, w, pv, mask, B, T, heads, d)
        close(kqv, kqv_ref, 2e-2, 2e-2, msg="kqv written by the fused kernel")      # bf16 rounding of two fp32 summation orders
        close(sc, sc_ref, 2e-2, 6e-2, msg="scores")
        close(out, out_ref, 3e-2, 3e-2, msg="attention output")
        # against the fp64 reference arithmetic (realformer.py:33-44) on the bf16-rounded operands
        kq = (x.double().view(B, T, heads, d) @ w.double().t()).view(B, T, heads, 3 * d)
        k_, q_, v_ = kq[..., :d], kq[..., d:2 * d], kq[..., 2 * d:]
        s_ = torch.einsum("bihd,bjhd->bhij", q_, k_) / math.sqrt(d)
        if pv is not None:
            s_ = s_ + pv.double()
        s_ = s_ - 10000.0 * (1.0 - mask.double())[:, None, :, None]
        o_ = torch.einsum("bhij,bjhd->bihd", torch.softmax(s_, -1), v_).reshape(B * T, H)
        close(out, o_, 5e-2, 5e-2, msg="vs fp64")


@pytest.mark.parametrize("B,T,heads,d", [(3, 28, 8, 96), (2, 75, 8, 96), (2, 64, 4, 64), (2, 10, 8, 16)])
def test_rf_attention_backward_fused_kqv_dgrad(B, T, heads, d):
    """attention backward with dx = dkqv . Wkqv + dres inside == attention backward + the separate dgrad GEMM."""
    bf = torch.bfloat16
    H = heads * d
    kqv = rnd(B * T * heads, 3 * d, seed=80).to(bf).to(DEV)
    w = (rnd(3 * d, d, seed=81) / math.sqrt(d)).to(bf).to(DEV)
    prev = rnd(B, heads, T, T, seed=82).to(DEV)
    mask = torch.ones(B, T, device=DEV)
    mask[0, T - T // 3:] = 0
    dout = rnd(B * T, H, seed=83).to(bf).to(DEV)
    dsc = (0.1 * rnd(B, heads, T, T, seed=84)).to(DEV)
    dres = rnd(B * T, H, seed=85).to(bf).to(DEV)
    _, scores = ops.rf_attn_fwd(kqv, prev, mask, B, T, heads, d)
    for ds_in, res in ((dsc, dres), (None, None)):
        dkqv_ref, dprev_ref = ops.rf_attn_bwd(kqv, scores, dout, ds_in, True, B, T, heads, d)
        dx_ref = torch.empty(B * T, H, device=DEV, dtype=bf)
        if res is not None:
            ops.gemm(B * T * heads, d, 3 * d, dkqv_ref, 3 * d, False, w, d, True, dx_ref, d, epilogue=EPI_RESIDUAL, aux_in=res,
                     ld_aux_in=d)
        else:
            ops.gemm(B * T * heads, d, 3 * d, dkqv_ref, 3 * d, False, w, d, True, dx_ref, d)
        dkqv, dprev, dx = ops.rf_attn_bwd_fused(kqv, scores, dout, ds_in, True, w, res, B, T, heads, d)
        assert torch.equal(dkqv, dkqv_ref) and torch.equal(dprev, dprev_ref)           # same code path for these two
        scale = dx_ref.float().abs().max().item()
        close(dx, dx_ref, 2e-2, 2e-2 * scale, msg="dx from the fused kernel")
        dx64 = (dkqv_ref.double() @ w.double()).view(B * T, H) + (res.double() if res is not None else 0.0)
        close(dx, dx64, 2e-2, 2e-2 * scale, msg="dx vs fp64")
