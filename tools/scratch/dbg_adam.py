import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn as nn
import mmvqa_b200
from mmvqa_b200 import functional as Fn
from mmvqa_b200.graph import GraphedTrainStep
from mmvqa_b200.models.realformer import ResEncoderBlock, run_blocks
from mmvqa_b200.optim import FusedAdam
DEV="cuda"
mmvqa_b200.set_compute_dtype(torch.float32)
xs = [torch.randn(4, 12, 128, device=DEV) for _ in range(3)]
mask = torch.ones(4, 12, device=DEV, dtype=torch.long)
def make():
    torch.manual_seed(3)
    return nn.ModuleList([ResEncoderBlock(emb_s=16, head_cnt=8, dp1=0.0, dp2=0.0) for _ in range(2)]).to(DEV)
def run(mode):
    blocks = make()
    opt = (torch.optim.Adam if mode == "torch" else FusedAdam)(blocks.parameters(), lr=1e-3)
    def loss_fn(x):
        h, _ = run_blocks(list(blocks), x, None, mask, False)
        return h.float().pow(2).mean()
    losses = []
    if mode == "graph":
        gs = GraphedTrainStep(loss_fn, [xs[0]], opt, warmup=0)
        for x in xs:
            losses.append(float(gs.replay(x))); print(mode, "step_dev", opt._step_dev.item())
    else:
        for x in xs:
            opt.zero_grad(set_to_none=True)
            loss = loss_fn(x); loss.backward(); opt.step(); losses.append(float(loss))
            if mode == "eager": print(mode, "step_dev", opt._step_dev.item())
    return losses
for m in ("torch", "eager", "graph"):
    print(m, run(m))
