"""Bring-up probe for the tcgen05 GEMM: every (major, tile) variant in its own process, exact integer
operands, bounded by a timeout.  Usage (GPU box):  python tools/tc_probe.py > gpurun_out/tc_probe.log"""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    # M, N, K, a_trans, b_trans
    (128, 128, 64, 0, 0), (128, 128, 256, 0, 0), (256, 512, 128, 0, 0), (128, 64, 64, 0, 0), (128, 32, 64, 0, 0),
    (128, 128, 64, 0, 1), (128, 256, 128, 0, 1), (128, 64, 64, 0, 1),
    (128, 128, 64, 1, 0), (256, 128, 128, 1, 0),
    (128, 128, 64, 1, 1), (256, 256, 192, 1, 1),
    (448, 3072, 768, 0, 0), (448, 768, 3072, 0, 0), (3584, 288, 96, 0, 0), (100, 72, 40, 0, 0), (8, 8, 8, 0, 0),
    (448, 768, 768, 0, 1), (768, 768, 448, 1, 1), (288, 96, 3584, 1, 1), (3584, 96, 288, 0, 1),
]

CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from mmvqa_b200 import ops
M, N, K, at, bt = %d, %d, %d, %d, %d
g = torch.Generator().manual_seed(1)
A = torch.randint(-2, 3, (M, K), generator=g).float()
B = torch.randint(-2, 3, (N, K), generator=g).float()
ref = A @ B.t()
As = (A.t().contiguous() if at else A).bfloat16().cuda()
Bs = (B.t().contiguous() if bt else B).bfloat16().cuda()
C = torch.full((M, N), -777.0, device="cuda")
ops.gemm(M, N, K, As, M if at else K, bool(at), Bs, N if bt else K, bool(bt), C, N)
torch.cuda.synchronize()
C = C.cpu()
bad = (C != ref)
if bad.any():
    idx = bad.nonzero()[:6].tolist()
    rows_bad = bad.any(1).nonzero().flatten().tolist()
    cols_bad = bad.any(0).nonzero().flatten().tolist()
    print("FAIL frac_bad=%%.4f maxerr=%%.1f untouched=%%d first=%%s" %% (bad.float().mean().item(), (C - ref).abs().max().item(),
          int((C == -777.0).sum()), [(i, j, C[i, j].item(), ref[i, j].item()) for i, j in idx]))
    print("  bad rows [%%d]: %%s ..." %% (len(rows_bad), rows_bad[:16]), " bad cols [%%d]: %%s ..." %% (len(cols_bad), cols_bad[:16]))
    # which 16-wide k slices explain the result?  least squares over per-slice partial products
    if K <= 256:
        parts = torch.stack([A[:, k:k + 16] @ B[:, k:k + 16].t() for k in range(0, K, 16)], 0).reshape(-1, M * N).t()
        sol = torch.linalg.lstsq(parts, C.reshape(-1, 1)).solution.flatten()
        print("  k-slice weights:", [round(float(x), 2) for x in sol])
else:
    print("PASS")
'''

if __name__ == "__main__":
    for case in CASES:
        code = CHILD % ((ROOT,) + case)
        try:
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
            out = (r.stdout.strip() or "") + ("" if r.returncode == 0 else "\n  rc=%d %s" % (r.returncode, r.stderr.strip()[-600:]))
        except subprocess.TimeoutExpired:
            out = "TIMEOUT"
        print(case, out, flush=True)
