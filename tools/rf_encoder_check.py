"""Parity + timing of the one-launch RealFormer encoder forward (csrc/rf_encoder.cu) against the per-operator chain
(same library, same bf16 arithmetic):  python tools/rf_encoder_check.py [--B 16] [--T 28] [--L 12] [--drop 0]

Compares every tensor RealFormerEncoderFn saves for the backward pass, the output and the parameter gradients of a
full forward+backward, then times both forwards with CUDA events."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--T", type=int, default=28)
    ap.add_argument("--L", type=int, default=12)
    ap.add_argument("--drop", type=float, default=0.0)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--trace", type=int, default=0)
    a = ap.parse_args()
    import mmvqa_b200
    from mmvqa_b200 import functional as Fn
    from mmvqa_b200.models.realformer import ResEncoderBlock, block_params
    mmvqa_b200.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(0)
    blocks = nn.ModuleList([ResEncoderBlock(emb_s=96, head_cnt=8, dp1=a.drop, dp2=a.drop) for _ in range(a.L)]).cuda()
    for b in blocks:            # LayerNorm / bias parameters away from their trivial initial values
        for p_ in (b.ln1.weight, b.ln2.weight):
            p_.data.uniform_(0.5, 1.5)
        for p_ in (b.ln1.bias, b.ln2.bias, b.ff[0].bias, b.ff[2].bias):
            p_.data.uniform_(-0.3, 0.3)
    params = []
    for b in blocks:
        params.extend(block_params(b))
    B, T = a.B, a.T
    x = torch.randn(B, T, 768, device="cuda").bfloat16()
    mask = torch.ones(B, T, device="cuda")
    for i in range(B):
        mask[i, T - (i % 5):] = 0.0
    prev0 = None

    def run(native: bool, xin):
        os.environ["MMVQA_RF_ENCODER"] = "1" if native else "0"
        seen = {}
        orig = torch.autograd.function.FunctionCtx.save_for_backward

        def spy(ctx, *ts):
            seen["saved"] = ts
            return orig(ctx, *ts)
        torch.autograd.function.FunctionCtx.save_for_backward = spy
        try:
            y, sc = Fn.RealFormerEncoderFn.apply(xin, mask, prev0, 8, a.drop, a.drop, 1234, *params)
        finally:
            torch.autograd.function.FunctionCtx.save_for_backward = orig
        return y, sc, seen["saved"]

    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    ya, sa, sva = run(False, xa)
    torch.cuda.synchronize()
    yb, sb, svb = run(True, xb)
    torch.cuda.synchronize()
    names = ["xin", "kqv", "scores", "attn", "y1", "mean1", "rstd1", "x1", "hpre", "hact", "y2", "mean2", "rstd2"]
    worst = 0.0
    bad = 0
    for l in range(a.L):
        for k, nm in enumerate(names):
            ta, tb = sva[l * 13 + k].float(), svb[l * 13 + k].float()
            if nm == "scores":          # masked query rows hold -10000 * layer: compare relative to that scale
                scale = max(1.0, ta.abs().max().item())
            else:
                scale = max(1e-3, ta.abs().max().item())
            if not torch.isfinite(tb).all():
                print("layer %d %s: NON-FINITE values in the cluster kernel output" % (l, nm))
                bad += 1
                continue
            err = (ta - tb).abs().max().item() / scale
            worst = max(worst, err)
            if l in (0, a.L - 1) or err > 3e-2:
                print("layer %2d %-6s max|diff|/max|ref| = %.3e" % (l, nm, err))
            if err > 6e-2:
                bad += 1
    erry = ((ya.float() - yb.float()).abs().max() / ya.float().abs().max()).item()
    print("output: max|diff|/max|ref| = %.3e   worst intermediate %.3e   failures %d" % (erry, worst, bad))
    # backward through both (the backward pass is the per-operator one either way: it consumes the saved tensors)
    g = torch.randn_like(ya.float()).bfloat16()
    for p_ in params:
        p_.grad = None
    ya.backward(g)
    ga = [p_.grad.clone() for p_ in params] + [xa.grad.clone()]
    for p_ in params:
        p_.grad = None
    yb.backward(g)
    gb = [p_.grad.clone() for p_ in params] + [xb.grad.clone()]
    gworst = 0.0
    for u, v in zip(ga, gb):
        gworst = max(gworst, ((u.float() - v.float()).abs().max() / max(1e-6, u.float().abs().max().item())).item())
    print("gradients (params + input): worst max|diff|/max|ref| = %.3e" % gworst)
    ok = bad == 0 and erry < 5e-2 and (a.drop > 0 or gworst < 1.5e-1)

    def timeit(native):
        with torch.no_grad():
            for _ in range(3):
                run(native, x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # eager launches are host bound for the per-operator chain: capture both in a CUDA graph
            g_ = torch.cuda.CUDAGraph()
            s_ = torch.cuda.Stream()
            s_.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_):
                run(native, x)
            torch.cuda.current_stream().wait_stream(s_)
            with torch.cuda.graph(g_):
                run(native, x)
            g_.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(a.reps):
                g_.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / a.reps * 1e3
    t_ops = timeit(False)
    t_one = timeit(True)
    print("forward, B=%d T=%d L=%d: per-operator chain %.1f us, cluster kernel %.1f us (%.2fx)" % (B, T, a.L, t_ops, t_one, t_ops / t_one))
    os.environ.pop("MMVQA_RF_ENCODER", None)
    if a.trace:
        from mmvqa_b200 import ops
        buf = torch.zeros(4, 16, 16, dtype=torch.int64, device="cuda")
        ops.RF_TRACE_BUFFER = buf
        with torch.no_grad():
            run(True, x)
        torch.cuda.synchronize()
        ops.RF_TRACE_BUFFER = None
        tr = buf.cpu()
        t0 = int(tr[3, 0, 0])
        ghz = 1.9
        roles = ["producerA", "producerB", "mma", "compute"]
        labels = {0: ["start", "proj issued", "ff1 issued", "ff2 issued"],
                  1: ["start", "B free (Wkqv)", "S1 seen", "S2 seen", "B free (x1)", "S3 seen", "B free (ff2)", "ff2 chunks issued"],
                  2: ["start", "att landed", "proj issued", "x1 landed", "ff1 t0", "ff1 t1", "ff1 t2", "ff2 first B", "ff2 issued"],
                  3: ["start", "wkqv landed", "kqv done", "attn done+S1", "proj acc", "y1 done", "LN1 stats", "x1 stored+S2",
                      "ff1 acc0", "ff1 acc1", "ff1 acc2", "hact stored+S3", "ff2 acc", "y2 done", "LN2 stats", "layer end"]}
        for l in range(min(a.L, a.trace)):
            print("---- layer %d (us since compute start of layer 0, CTA 0; clock64 / %.1f GHz)" % (l, ghz))
            for r in range(4):
                print("  %-9s " % roles[r] + "  ".join("%s=%.1f" % (labels[r][k], (int(tr[r, l, k]) - t0) / (ghz * 1e3))
                                                     for k in range(len(labels[r]))))
    print("RESULT", "OK" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
