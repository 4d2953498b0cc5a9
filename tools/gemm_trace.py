"""Phase timestamps (%globaltimer) of the tcgen05 GEMM kernel at the flagship shapes: where do the ~10 us of a
448-row GEMM go?   python tools/gemm_trace.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mmvqa_b200 import ops  # noqa: E402
from mmvqa_b200._lib import ACT_SERF, EPI_ACT, EPI_RESIDUAL  # noqa: E402

bf = torch.bfloat16
M, H, F4 = 448, 768, 3072
NAMES = ["entry", "setup", "depwait", "loads_issued", "first_tile", "last_mma_issued", "acc_ready", "epi_done", "exit"]


def r(*s):
    return (torch.randn(*s, device="cuda") * 0.5).to(bf)


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(name, Mx, N, K, bt=False, **kw):
    A, Bm = r(Mx, K), (r(K, N) if bt else r(N, K))
    cdt = kw.pop("cdt", bf)
    ns = kw.get("split_k", 1)
    C = torch.empty(ns, Mx, N, device="cuda", dtype=cdt)
    trace = torch.zeros(4096, 16, dtype=torch.int64, device="cuda")
    for it in range(3):
        flush.zero_()
        trace.zero_()
        torch.cuda.synchronize()
        ops.gemm(Mx, N, K, A, K, False, Bm, N if bt else K, bt, C, N, trace=trace, **kw)
        torch.cuda.synchronize()
    t = trace.cpu()
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t[:, :9] - t0).float() / 1e3      # us
    print(f"{name}: {t.shape[0]} CTAs; kernel span (first entry -> last exit) {rel[:, 8].max():.2f} us")
    print("   phase            min     median  max   (us after the first CTA entry)")
    for i, n in enumerate(NAMES):
        col = rel[:, i]
        print(f"   {n:16s} {col.min():6.2f} {col.median():7.2f} {col.max():6.2f}")
    rel2 = (t[:, 9:11] - t0).float() / 1e3
    print(f"   epilogue detail (median): acc_ready {rel[:, 6].median():.2f} -> first tcgen05.ld done {rel2[:, 0].median():.2f} -> "
          f"first chunk stored {rel2[:, 1].median():.2f} -> epi_done {rel[:, 7].median():.2f}")
    d = rel[:, 1:9] - rel[:, 0:8]
    print("   per-CTA deltas (median): " + "  ".join(f"{NAMES[i + 1]}+{d[:, i].median():.2f}" for i in range(8)))


bias_f, bias_h = torch.randn(F4, device="cuda"), torch.randn(H, device="cuda")
pre = torch.empty(M, F4, device="cuda", dtype=bf)
run("ff1 plain", M, F4, H)
run("ff1 +bias+SERF", M, F4, H, bias=bias_f, epilogue=EPI_ACT, act=ACT_SERF, aux_out=pre, ld_aux_out=F4)
run("ff2 +bias+residual", M, H, F4, bias=bias_h, epilogue=EPI_RESIDUAL, aux_in=r(M, H), ld_aux_in=H)
run("ff2 split3 slabs", M, H, F4, bias=bias_h, split_k=3, c_split_stride=M * H, cdt=torch.float32)
run("ff1 dgrad split3 slabs (B MN-major)", M, H, F4, bt=True, split_k=3, c_split_stride=M * H, cdt=torch.float32)
run("proj +residual", M, H, H, epilogue=EPI_RESIDUAL, aux_in=r(M, H), ld_aux_in=H)
run("kqv", M * 8, 288, 96)
if "--more" in sys.argv:
    run("ff2 split4 slabs", M, H, F4, bias=bias_h, split_k=4, c_split_stride=M * H, cdt=torch.float32)
    run("ff2 split6 slabs", M, H, F4, bias=bias_h, split_k=6, c_split_stride=M * H, cdt=torch.float32)
    run("ff2 split8 slabs", M, H, F4, bias=bias_h, split_k=8, c_split_stride=M * H, cdt=torch.float32)
if "--big" in sys.argv:
    run("mid 4096x3072x768 plain", 4096, F4, H)
    run("mid 4096x3072x768 +bias+SERF", 4096, F4, H, bias=bias_f, epilogue=EPI_ACT, act=ACT_SERF,
        aux_out=torch.empty(4096, F4, device="cuda", dtype=bf), ld_aux_out=F4)
    run("mid 4096x768x3072 +bias+residual", 4096, H, F4, bias=bias_h, epilogue=EPI_RESIDUAL, aux_in=r(4096, H), ld_aux_in=H)
    run("wgrad-like 3072x768x4096 fp32 out", F4, H, 4096, cdt=torch.float32)
