"""one mid-size GEMM launch (4096 x 3072 x 768, bf16 plain store) for `ncu --set full` (stall reasons of the epilogue)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mmvqa_b200 import ops  # noqa: E402

bf = torch.bfloat16
a = (torch.randn(4096, 768, device="cuda") * 0.5).to(bf)
b = (torch.randn(3072, 768, device="cuda") * 0.5).to(bf)
c = torch.empty(4096, 3072, device="cuda", dtype=bf)
for _ in range(3):
    ops.gemm(4096, 3072, 768, a, 768, False, b, 768, False, c, 3072)
torch.cuda.synchronize()
print("done")
