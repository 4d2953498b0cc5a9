"""BASELINE.json configs[4]: encoder-only sweep (fwd+bwd), seq len x batch x {RealFormer, Transformer}, bf16.
Prints samples/s and the fraction of the measured bf16 tensor-core peak (algorithmic GEMM FLOPs, step = 3 x fwd).
    python tools/sweep.py [--quick]"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

import mmvqa_b200  # noqa: E402
from mmvqa_b200.models.realformer import ResEncoderBlock, run_blocks  # noqa: E402
from mmvqa_b200.models.transformer import BertLayer  # noqa: E402

H, L = 768, 12
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1590.0


def flops_per_sample(kind, T):
    per_tok = (24 * H * H + 4 * T * H) if kind == "transformer" else ((0.75 + 2 + 16) * H * H + 4 * T * H)
    return 3 * L * T * per_tok


def bench(kind, B, T, iters=5):
    torch.manual_seed(0)
    if kind == "realformer":
        blocks = [ResEncoderBlock(emb_s=96, head_cnt=8, dp1=0.1, dp2=0.1).cuda().train() for _ in range(L)]
        params = [p for b in blocks for p in b.parameters()]

        def step(x, mask):
            y, _ = run_blocks(blocks, x, None, mask, True)
            return y
    else:
        args = types.SimpleNamespace(hidden_size=H, heads=12, hidden_dropout_prob=0.1, n_layers=L)
        layer = BertLayer(args, share="none", norm="pre").cuda().train()
        params = list(layer.parameters())

        def step(x, mask):
            h = x
            for i in range(L):
                h = layer(h, mask, i)
            return h
    x = torch.randn(B, T, H, device="cuda").bfloat16().requires_grad_(True)
    mask = torch.ones(B, T, dtype=torch.long, device="cuda")
    mask[:, T - T // 4:] = 0
    go = torch.randn(B, T, H, device="cuda").bfloat16()

    def one():
        for p in params:
            p.grad = None
        x.grad = None
        step(x, mask).backward(go)
    for _ in range(2):
        one()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        one()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / iters
    sps = B / (ms * 1e-3)
    frac = sps * flops_per_sample(kind, T) / (peak * 1e12)
    del g
    return ms, sps, frac


EFFNET_MAPS = [(24, 112), (48, 56), (80, 28), (176, 14), (512, 7)]


def bench_projector(B, num_vis, act, iters=5):
    """visual-token projector forward + backward (image_encoding.py:100-115) over the first `num_vis` pyramid levels
    (num_vis < 5 = "first k tokens of the 5"); act = projector activation (SERF default, ReLU with
    --use_relu, image_encoding.py:68,94)."""
    from mmvqa_b200 import functional as Fn
    from mmvqa_b200._lib import ACT_RELU, ACT_SERF
    torch.manual_seed(0)
    feats = [torch.randn(B, c, s, s, device="cuda").abs_().requires_grad_(True) for c, s in EFFNET_MAPS[:num_vis]]
    ws = [torch.randn(H, c, 1, 1, device="cuda").mul_(c ** -0.5).requires_grad_(True) for c, _ in EFFNET_MAPS[:num_vis]]
    go = torch.randn(num_vis, B, H, device="cuda")
    code = ACT_SERF if act == "serf" else ACT_RELU

    def one():
        for t in feats + ws:
            t.grad = None
        Fn.vistok_project_all(feats, ws, code).backward(go)
    for _ in range(2):
        one()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        one()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 3 * sum(2 * s * s * c * H for c, s in EFFNET_MAPS[:num_vis])
    sps = B / (ms * 1e-3)
    del g
    return ms, sps, sps * flops / (peak * 1e12)


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    mmvqa_b200.set_compute_dtype("bf16")
    Ts = [28, 128] if quick else [28, 32, 64, 75, 96, 128]
    Bs = [16, 256, 1024] if quick else [16, 64, 256, 1024]
    print(f"encoder-only fwd+bwd, bf16, 12 layers, hidden 768; peak {peak} TFLOP/s (measured)")
    for kind in ("realformer", "transformer"):
        for T in Ts:
            for B in Bs:
                try:
                    ms, sps, frac = bench(kind, B, T)
                    print(f"{kind:11s} T={T:4d} B={B:5d}  {ms:9.3f} ms/step  {sps:10.0f} samples/s  {100 * frac:5.1f}% of bf16 peak",
                          flush=True)
                except Exception as e:   # noqa: BLE001
                    print(f"{kind:11s} T={T:4d} B={B:5d}  FAILED {type(e).__name__}: {str(e)[:120]}", flush=True)
                torch.cuda.empty_cache()
    print("visual-token projector fwd+bwd (EfficientNetV2-M map shapes), bf16")
    for act in ("serf", "relu"):
        for nv in ([1, 5] if quick else [1, 2, 3, 4, 5]):
            for B in ([16] if quick else [16, 64]):
                try:
                    ms, sps, frac = bench_projector(B, nv, act)
                    print(f"projector {act:4s} num_vis={nv} B={B:4d}  {ms:9.3f} ms/step  {sps:10.0f} samples/s  {100 * frac:5.1f}% of bf16 peak",
                          flush=True)
                except Exception as e:   # noqa: BLE001
                    print(f"projector {act:4s} num_vis={nv} B={B:4d}  FAILED {type(e).__name__}: {str(e)[:120]}", flush=True)
                torch.cuda.empty_cache()
