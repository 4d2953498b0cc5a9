"""A handful of launches whose tensor-pipe utilisation the VERDICT asks for (ncu --metrics ..., see tools/ncu_tensor.sh):
GEMM 8192^3, 4096x3072x768, the M = 448 flagship GEMMs, RealFormer attention fwd / bwd, the one-launch encoder."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mmvqa_b200  # noqa: E402
from mmvqa_b200 import ops  # noqa: E402
from mmvqa_b200._lib import ACT_SERF, EPI_ACT  # noqa: E402

bf = torch.bfloat16
mmvqa_b200.set_compute_dtype(bf)


def r(*s):
    return (torch.randn(*s, device="cuda") * 0.5).to(bf)


def gemm(M, N, K, **kw):
    a, b = r(M, K), r(N, K)
    c = torch.empty(M, N, device="cuda", dtype=bf)
    for _ in range(2):
        ops.gemm(M, N, K, a, K, False, b, K, False, c, N, **kw)
    torch.cuda.synchronize()


gemm(8192, 8192, 8192)
gemm(4096, 3072, 768)
M = 448
pre = torch.empty(M, 3072, device="cuda", dtype=bf)
gemm(M, 3072, 768, bias=torch.randn(3072, device="cuda"), epilogue=EPI_ACT, act=ACT_SERF, aux_out=pre, ld_aux_out=3072)
gemm(M, 768, 768)
B, T, heads, d = 16, 28, 8, 96
kqv = r(M * heads, 3 * d)
mask = torch.ones(B, T, device="cuda")
prev = torch.randn(B, heads, T, T, device="cuda")
do = r(M, 768)
for _ in range(2):
    out, sc = ops.rf_attn_fwd(kqv, prev, mask, B, T, heads, d)
    ops.rf_attn_bwd(kqv, sc, do, prev, True, B, T, heads, d)
torch.cuda.synchronize()
print("probe done")
