"""Standalone timings of the hot kernels at the flagship shapes (GPU box).  Graph-captured back-to-back launches,
CUDA-event timed; 'cold' = 256 MB L2 flush before every launch, 'warm' = no flush.
    python tools/kernel_bench.py [B] [T]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mmvqa_b200 import ops  # noqa: E402
from mmvqa_b200._lib import ACT_SERF, EPI_ACT, EPI_ACT_ROWSUM, EPI_DACT, EPI_DACT_SCALE, EPI_RESIDUAL  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 28
M, H, F4, heads, d = B * T, 768, 3072, 8, 96
bf = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20, cold=True):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()

    def cap(with_fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                if cold:
                    flush.zero_()
                if with_fn:
                    fn()
        return g

    def run(g):
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)
    t = run(cap(True))
    if cold:
        t -= run(cap(False))
    return t / iters * 1e3   # us


def r(*s, dt=bf):
    return (torch.randn(*s, device="cuda") * 0.5).to(dt)


def report(name, fn, flops=0, bytes_=0):
    tc, tw = timeit(fn, cold=True), timeit(fn, cold=False)
    extra = ""
    if flops:
        extra += f"  {flops / tc / 1e6:8.1f} TF/s cold {flops / tw / 1e6:8.1f} warm"
    if bytes_:
        extra += f"  {bytes_ / tc / 1e3:8.1f} GB/s cold {bytes_ / tw / 1e3:8.1f} warm"
    print(f"{name:58s} cold {tc:8.2f} us  warm {tw:8.2f} us{extra}", flush=True)


def gemm_case(name, Mx, N, K, at=False, bt=False, cdt=bf, **kw):
    A = r(K, Mx) if at else r(Mx, K)
    Bm = r(K, N) if bt else r(N, K)
    C = torch.zeros(Mx, N, device="cuda", dtype=cdt)
    report(name + f" [{Mx}x{N}x{K} {'T' if at else 'N'}{'T' if bt else 'N'}]",
           lambda: ops.gemm(Mx, N, K, A, Mx if at else K, at, Bm, N if bt else K, bt, C, N, **kw), flops=2.0 * Mx * N * K)


ATTN_ONLY = os.environ.get("KB_ATTN_ONLY") is not None
print(f"B={B} T={T} M={M}  env BN={os.environ.get('MMVQA_TC_BN')} STAGES={os.environ.get('MMVQA_TC_STAGES')}")
bias_h, bias_f = torch.randn(H, device="cuda"), torch.randn(F4, device="cuda")
if ATTN_ONLY:      # KB_ATTN_ONLY=1 python tools/kernel_bench.py 16 75: the attention rows at another sequence length
    kqv = r(M * heads, 3 * d)
    mask = torch.ones(B, T, device="cuda")
    prev = torch.randn(B, heads, T, T, device="cuda")
    xin, wk = r(M, H), r(3 * d, d)
    report(f"rf_attn_fwd T={T}", lambda: ops.rf_attn_fwd(kqv, prev, mask, B, T, heads, d), flops=4.0 * B * heads * T * T * d)
    report(f"rf_attn_fwd_fused (kqv inside) T={T}", lambda: ops.rf_attn_fwd_fused(xin, wk, prev, mask, B, T, heads, d),
           flops=4.0 * B * heads * T * T * d + 2.0 * M * heads * 3 * d * d)
    out, sc = ops.rf_attn_fwd(kqv, prev, mask, B, T, heads, d)
    do = r(M, H)
    report(f"rf_attn_bwd T={T}", lambda: ops.rf_attn_bwd(kqv, sc, do, prev, True, B, T, heads, d), flops=8.0 * B * heads * T * T * d)
    sys.exit(0)
gemm_case("kqv fwd", M * heads, 3 * d, d)
gemm_case("proj fwd +residual", M, H, H, epilogue=EPI_RESIDUAL, aux_in=r(M, H), ld_aux_in=H)
pre = torch.empty(M, F4, device="cuda", dtype=bf)
gemm_case("ff1 fwd +bias+SERF", M, F4, H, bias=bias_f, epilogue=EPI_ACT, act=ACT_SERF, aux_out=pre, ld_aux_out=F4)
gemm_case("ff1 fwd plain", M, F4, H)
gemm_case("ff2 fwd +bias+residual", M, H, F4, bias=bias_h, epilogue=EPI_RESIDUAL, aux_in=r(M, H), ld_aux_in=H)
gemm_case("ff2 dgrad +dSERF", M, F4, H, bt=True, epilogue=EPI_DACT, act=ACT_SERF, aux_in=r(M, F4), ld_aux_in=F4)
gemm_case("ff1 dgrad +residual", M, H, F4, bt=True, epilogue=EPI_RESIDUAL, aux_in=r(M, H), ld_aux_in=H)
gemm_case("proj dgrad", M, H, H, bt=True)
gemm_case("kqv dgrad +residual", M * heads, d, 3 * d, bt=True, epilogue=EPI_RESIDUAL, aux_in=r(M * heads, d), ld_aux_in=d)
gemm_case("ff2 wgrad", H, F4, M, at=True, bt=True, cdt=torch.float32)
gemm_case("ff1 wgrad", F4, H, M, at=True, bt=True, cdt=torch.float32)
gemm_case("proj wgrad", H, H, M, at=True, bt=True, cdt=torch.float32)
gemm_case("proj wgrad split4", H, H, M, at=True, bt=True, cdt=torch.float32, accumulate=True, split_k=4)
gemm_case("kqv wgrad", 3 * d, d, M * heads, at=True, bt=True, cdt=torch.float32)
gemm_case("kqv wgrad split14", 3 * d, d, M * heads, at=True, bt=True, cdt=torch.float32, accumulate=True, split_k=14)
for sk in (2, 3, 4, 6, 8):
    gemm_case(f"ff2 fwd split{sk} (fp32 atomics)", M, H, F4, cdt=torch.float32, accumulate=True, split_k=sk)
for sk in (2, 3, 4):
    gemm_case(f"proj fwd split{sk} (fp32 atomics)", M, H, H, cdt=torch.float32, accumulate=True, split_k=sk)
for sk in (2, 3):
    gemm_case(f"ff1 fwd split{sk} (fp32 atomics)", M, F4, H, cdt=torch.float32, accumulate=True, split_k=sk)
if os.environ.get("KB_ONLY_SPLIT"):
    sys.exit(0)
gemm_case("big 8192^3", 8192, 8192, 8192)
gemm_case("big 4096x3072x768", 4096, 3072, 768)

# projector levels (EffNetV2-M maps), forward pooled epilogue and backward recompute
for Cc, side in [(24, 112), (48, 56), (80, 28), (176, 14), (512, 7)]:
    HW = side * side
    ld = (HW + 7) // 8 * 8
    W = r(H, Cc)
    f = r(B * Cc, ld).abs()
    v = torch.zeros(B, H, device="cuda")
    report(f"projector fwd C={Cc} HW={HW}", lambda: ops.gemm(H, HW, Cc, W, Cc, False, f, ld, True, None, 0, epilogue=EPI_ACT_ROWSUM,
           act=ACT_SERF, rowsum_out=v, scale=1.0 / HW, batch=B, a_batch_rows=0, b_batch_rows=Cc), flops=2.0 * B * H * HW * Cc)
    G = torch.empty(B, H, ld, device="cuda", dtype=bf)
    dv = torch.randn(B, H, device="cuda")
    report(f"projector bwd recompute C={Cc} HW={HW}", lambda: ops.gemm(H, HW, Cc, W, Cc, False, f, ld, True, G, ld,
           epilogue=EPI_DACT_SCALE, act=ACT_SERF, rowscale=dv, scale=1.0 / HW, batch=B, a_batch_rows=0, b_batch_rows=Cc,
           c_batch_stride=H * ld), flops=2.0 * B * H * HW * Cc)
    dw = torch.zeros(H, Cc, device="cuda")
    report(f"projector dW C={Cc} HW={HW}", lambda: ops.gemm(H, Cc, HW, G, ld, False, f, ld, False, dw, Cc, accumulate=True, batch=B,
           a_batch_rows=H, b_batch_rows=Cc, c_batch_stride=0), flops=2.0 * B * H * HW * Cc)

# attention, LN, elementwise
kqv = r(M * heads, 3 * d)
mask = torch.ones(B, T, device="cuda")
prev = torch.randn(B, heads, T, T, device="cuda")
report("rf_attn_fwd", lambda: ops.rf_attn_fwd(kqv, prev, mask, B, T, heads, d))
out, sc = ops.rf_attn_fwd(kqv, prev, mask, B, T, heads, d)
do = r(M, H)
report("rf_attn_bwd", lambda: ops.rf_attn_bwd(kqv, sc, do, prev, True, B, T, heads, d))
x, res = r(M, H), r(M, H)
g, b_ = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
report("add_layernorm_fwd", lambda: ops.add_layernorm_fwd(x, None, g, b_, 1e-5, False), bytes_=2 * M * H * 2)
y, _, mean, rstd = ops.add_layernorm_fwd(x, None, g, b_, 1e-5, False)
dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
report("layernorm_bwd", lambda: ops.layernorm_bwd(do, x, g, mean, rstd, None, dg, db), bytes_=3 * M * H * 2)
xb = r(M, F4)
report("bias_act_fwd SERF", lambda: ops.bias_act_fwd(xb, bias_f, ACT_SERF), bytes_=2 * M * F4 * 2)
report("colsum [M,3072]", lambda: ops.colsum(xb, M, F4), bytes_=M * F4 * 2)
report("dropout [M,768]", lambda: ops.dropout(x, 0.1, 1), bytes_=2 * M * H * 2)
w32 = torch.randn(F4, H, device="cuda")
report("cast weight f32->bf16 [3072,768]", lambda: ops.cast(w32, bf), bytes_=F4 * H * 6)
big = torch.randn(64 << 20, device="cuda")
report("cast f32->bf16 64M", lambda: ops.cast(big, bf), bytes_=(64 << 20) * 6)
xl = r(65536, H)
report("add_layernorm_fwd 65536 rows", lambda: ops.add_layernorm_fwd(xl, None, g, b_, 1e-5, False), bytes_=2 * 65536 * H * 2)
