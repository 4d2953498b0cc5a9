"""The latency-bound main-chain kernels of the RealFormer backward at the flagship shape, for `ncu --set full`:
attention backward, LayerNorm backward (plain and split-K-slab input), LayerNorm forward over slabs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mmvqa_b200  # noqa: E402
from mmvqa_b200 import ops  # noqa: E402

bf = torch.bfloat16
mmvqa_b200.set_compute_dtype(bf)
B, T, heads, d, H = 16, 28, 8, 96, 768
M = B * T


def r(*s):
    return (torch.randn(*s, device="cuda") * 0.5).to(bf)


kqv = r(M * heads, 3 * d)
mask = torch.ones(B, T, device="cuda")
prev = torch.randn(B, heads, T, T, device="cuda")
do = r(M, H)
x = r(M, H)
g, b_ = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
y, _, mean, rstd = ops.add_layernorm_fwd(x, None, g, b_, 1e-5, False)
dg, db, ds = (torch.zeros(H, device="cuda") for _ in range(3))
parts = torch.randn(4, M, H, device="cuda")
for _ in range(3):
    out, sc = ops.rf_attn_fwd(kqv, prev, mask, B, T, heads, d)
    ops.rf_attn_bwd(kqv, sc, do, prev, True, B, T, heads, d)
    ops.layernorm_bwd(do, x, g, mean, rstd, None, dg, db, want_drop=True, dxsum=ds, dropout_p=0.1, dropout_seed=3)
    ops.layernorm_bwd_parts(parts, do, x, g, mean, rstd, dg, db, want_drop=True, dropout_p=0.1, dropout_seed=3)
    ops.add_layernorm_fwd_parts(parts, x, g, b_, 1e-5, bf, 0.1, 5)
torch.cuda.synchronize()
print("probe done")
