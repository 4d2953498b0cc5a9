"""Row-gated Adam on a bert-base word-embedding table: time of one launch with 448 live rows vs the dense update."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mmvqa_b200.optim import FusedAdam  # noqa: E402

V, H = 30522, 768


def run(gated):
    w = torch.randn(V, H, device="cuda").requires_grad_(True)
    opt = FusedAdam([w], lr=1e-3)
    ids = torch.randint(1000, V, (16, 28), device="cuda")
    if gated:
        opt.register_row_sparse(w, lambda: ids)
    w.grad = torch.zeros_like(w)
    w.grad.index_add_(0, ids.reshape(-1), torch.randn(ids.numel(), H, device="cuda"))
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        opt.step()
    e1.record()
    e1.synchronize()
    print("gated" if gated else "dense", "%.1f us per step" % (e0.elapsed_time(e1) / 20 * 1e3))


run(False)
run(True)
