"""World-size-2 gloo tests (CPU) of the data-parallel host logic: bucketed gradient all-reduce == single-process
gradients on the concatenated batch, and the SupCon all-gather (forward order + backward reduction) == the oracle
on the global batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from oracle import mmbert_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmvqa_b200.parallel import (GradBuckets, LayerwiseReducer, broadcast_parameters, gather_mask_rows,
                                     gather_supcon_features)
    torch.manual_seed(100 + rank)                      # different init per rank on purpose
    model = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 3))
    broadcast_parameters(model)
    g = torch.Generator().manual_seed(7)
    X, Y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]
    loss = ((model(xs) - ys) ** 2).mean()
    loss.backward()
    buckets = GradBuckets(list(model.parameters()), bucket_mb=1e-4)       # tiny buckets: several all-reduces
    assert len(buckets.buckets) > 1
    buckets.reduce()
    avg = [gv * buckets.grad_scale for gv in buckets.grads()]
    # layer-granular exchange (the backward gradient sink calls it once per layer, step() once for the rest)
    red = LayerwiseReducer(torch.float32)
    plist = list(model.parameters())
    lay = []
    for grp in (plist[2:], plist[:2]):                                    # last layer first, as in backward
        lay = [v.clone() * red.grad_scale for v in red(grp, [p.grad for p in grp])] + lay
    again = red(plist[2:], [p.grad for p in plist[2:]])                   # second step reuses the same bucket
    assert len(red._buckets) == 2 and again[0].data_ptr() == red._buckets[(id(plist[2]), 2)][1][0].data_ptr()
    # SupCon: each rank holds 3 samples x 2 views; gathered batch = 6 samples
    F = torch.randn(6, 2, 8, generator=g)
    F = F / F.norm(dim=-1, keepdim=True)
    soft = torch.rand(6, 6, generator=g)
    soft.fill_diagonal_(1.0)
    f_local = F[rank * 3:(rank + 1) * 3].clone().requires_grad_(True)
    gathered = gather_supcon_features(f_local)
    full_mask = gather_mask_rows(soft[rank * 3:(rank + 1) * 3])
    sl = O.supcon_loss(gathered, mask=full_mask)
    sl.backward()
    if rank == 0:
        q.put({"state": {k: v.clone() for k, v in model.state_dict().items()}, "avg": [a.clone() for a in avg],
               "lay": lay,
               "gathered": gathered.detach().clone(), "supcon": sl.detach().clone(), "mask": full_mask.clone()})
    q.put({"rank": rank, "df": f_local.grad.clone()})
    dist.barrier()
    dist.destroy_process_group()


def test_dp_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(3)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    main = next(x for x in got if "state" in x)
    dfs = {x["rank"]: x["df"] for x in got if "rank" in x}
    # single-process reference on the concatenated batch
    model = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 3))
    model.load_state_dict(main["state"])
    g = torch.Generator().manual_seed(7)
    X, Y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    ((model(X) - Y) ** 2).mean().backward()
    for a, p in zip(main["avg"], model.parameters()):
        torch.testing.assert_close(a, p.grad, rtol=1e-5, atol=1e-6)
    for a, p in zip(main["lay"], model.parameters()):
        torch.testing.assert_close(a, p.grad, rtol=1e-5, atol=1e-6)
    F = torch.randn(6, 2, 8, generator=g)
    F = (F / F.norm(dim=-1, keepdim=True)).requires_grad_(True)
    soft = torch.rand(6, 6, generator=g)
    soft.fill_diagonal_(1.0)
    torch.testing.assert_close(main["gathered"], F.detach())
    torch.testing.assert_close(main["mask"], soft)
    ref = O.supcon_loss(F, mask=soft)
    torch.testing.assert_close(main["supcon"], ref.detach())
    ref.backward()
    # every rank back-propagates the same global loss; the gather's backward sums the contributions of all
    # ranks, so each local slice receives world x the single-process gradient (DP averaging divides it back)
    for r in range(2):
        torch.testing.assert_close(dfs[r] / 2, F.grad[r * 3:(r + 1) * 3], rtol=1e-5, atol=1e-6)
