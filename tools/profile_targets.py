"""Runs each hot kernel a few times at the flagship shapes so that `ncu --set full -k regex:...` can capture them."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mmvqa_b200 import ops  # noqa: E402
from mmvqa_b200._lib import ACT_SERF, EPI_ACT, EPI_ACT_ROWSUM, EPI_DACT_SCALE, EPI_RESIDUAL  # noqa: E402

B, T, H, F4, heads, d = 16, 28, 768, 3072, 8, 96
M = B * T
bf = torch.bfloat16


def r(*s, dt=bf):
    return (torch.randn(*s, device="cuda") * 0.5).to(dt)


kqv = r(M * heads, 3 * d)
mask = torch.ones(B, T, device="cuda")
prev = torch.randn(B, heads, T, T, device="cuda")
do = r(M, H)
x = r(M, H)
g, b_ = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
W0, f0 = r(H, 24), r(B * 24, 12544).abs()
v = torch.zeros(B, H, device="cuda")
G = torch.empty(B, H, 12544, device="cuda", dtype=bf)
dv = torch.randn(B, H, device="cuda")
xa, wa, ba = r(M, H), r(F4, H), torch.randn(F4, device="cuda")
ya, pa = torch.empty(M, F4, device="cuda", dtype=bf), torch.empty(M, F4, device="cuda", dtype=bf)
xb, wb, bb = r(M, F4), r(H, F4), torch.randn(H, device="cuda")
yb, rb = torch.empty(M, H, device="cuda", dtype=bf), r(M, H)
for _ in range(3):
    out, sc = ops.rf_attn_fwd(kqv, prev, mask, B, T, heads, d)
    ops.rf_attn_bwd(kqv, sc, do, prev, True, B, T, heads, d)
    y, _, mean, rstd = ops.add_layernorm_fwd(x, None, g, b_, 1e-5, False)
    dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    ops.layernorm_bwd(do, x, g, mean, rstd, None, dg, db)
    ops.gemm(H, 12544, 24, W0, 24, False, f0, 12544, True, None, 0, epilogue=EPI_ACT_ROWSUM, act=ACT_SERF, rowsum_out=v,
             scale=1.0 / 12544, batch=B, a_batch_rows=0, b_batch_rows=24)
    ops.gemm(H, 12544, 24, W0, 24, False, f0, 12544, True, G, 12544, epilogue=EPI_DACT_SCALE, act=ACT_SERF, rowscale=dv,
             scale=1.0 / 12544, batch=B, a_batch_rows=0, b_batch_rows=24, c_batch_stride=H * 12544)
    ops.gemm(M, F4, H, xa, H, False, wa, H, False, ya, F4, bias=ba, epilogue=EPI_ACT, act=ACT_SERF, aux_out=pa, ld_aux_out=F4)
    ops.gemm(M, H, F4, xb, F4, False, wb, F4, False, yb, H, bias=bb, epilogue=EPI_RESIDUAL, aux_in=rb, ld_aux_in=H)
torch.cuda.synchronize()
# round 2: projector forward with the weight-gradient contraction inside (112 x 112 level), multi-tensor cast
Wp = r(H, 24)
fmaps = [torch.randn(B * c, s * s, device="cuda").abs() for c, s in ((24, 112), (48, 56), (80, 28), (176, 14), (512, 7))]
for _ in range(3):
    vv = torch.zeros(B, H, device="cuda")
    ops.vistok_fwd_pgrad(Wp, f0, 12544, vv, B, H, 12544, 24, ACT_SERF)
    ops.cast_pad_multi(fmaps, [(t.shape[1] + 7) // 8 * 8 for t in fmaps])
torch.cuda.synchronize()
print("done")
