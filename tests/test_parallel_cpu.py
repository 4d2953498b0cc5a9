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


def _supcon_rows_torch(features, mask, temperature=0.07, base_temperature=0.07):
    """per-anchor SupCon losses (loss.py:58-96 before the final mean), contrast_mode 'all'."""
    bsz, nv, _ = features.shape
    cf = features.transpose(0, 1).reshape(nv * bsz, -1)
    logits = cf @ cf.t() / temperature
    logits = logits - logits.max(dim=1, keepdim=True)[0].detach()
    m = mask.float().repeat(nv, nv)
    lm = 1.0 - torch.eye(nv * bsz)
    m = m * lm
    exp_logits = torch.exp(logits) * lm
    log_prob = logits - torch.log(exp_logits.sum(1, keepdim=True))
    return -(temperature / base_temperature) * (m * log_prob).sum(1) / m.sum(1)


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
    # row-sparse exchange of an embedding-table gradient: all-gather of the touched rows == dense all-reduce(SUM)
    emb = nn.Parameter(torch.zeros(20, 4))
    other = nn.Parameter(torch.zeros(3))
    ids = torch.tensor([[1, 5, 5, 7], [0, 5, 19, 1]][rank])              # duplicates inside a rank and across ranks
    contrib = torch.randn(4, 4, generator=torch.Generator().manual_seed(50 + rank))
    g_emb = torch.zeros(20, 4).index_add_(0, ids, contrib)
    g_other = torch.full((3,), float(rank + 1))
    red2 = LayerwiseReducer(torch.float32)
    red2.register_row_sparse(emb, lambda: ids)
    out2 = red2([emb, other], [g_emb, g_other])
    dense_sum = g_emb.clone()
    dist.all_reduce(dense_sum)
    q.put({"rank3": rank, "sparse": out2[0].clone(), "dense": dense_sum, "other": out2[1].clone()})
    # a second step with other ids: the accumulator clears the rows of the previous exchange only (not the whole table)
    ids2 = torch.tensor([[2, 2, 9, 7], [3, 18, 18, 9]][rank])
    contrib2 = torch.randn(4, 4, generator=torch.Generator().manual_seed(60 + rank))
    g_emb2 = torch.zeros(20, 4).index_add_(0, ids2, contrib2)
    ids.copy_(ids2)                                                       # the registered ids tensor is a fixed buffer
    out3 = red2([emb, other], [g_emb2, g_other])
    dense_sum2 = g_emb2.clone()
    dist.all_reduce(dense_sum2)
    q.put({"rank3b": rank, "sparse": out3[0].clone(), "dense": dense_sum2,
           "gathered": sorted(red2.gathered_ids(emb).tolist())})
    # local-anchor rows: the same gather with a reduce-scatter backward; each rank evaluates only the
    # loss rows of its own samples' anchors (here with plain torch ops restating loss.py:72-96 per anchor row)
    from mmvqa_b200.parallel import _GatherFeaturesRS
    f2 = F[rank * 3:(rank + 1) * 3].clone().requires_grad_(True)
    g2 = _GatherFeaturesRS.apply(f2)
    rows = _supcon_rows_torch(g2, full_mask)                              # [2 * 6] per-anchor losses, view-major
    mine = torch.cat([rows[v * 6 + rank * 3: v * 6 + rank * 3 + 3] for v in range(2)])
    loss_r = mine.sum() * (2.0 / 12.0)                                   # world * sum(local rows) / (n_views * bsz_global)
    loss_r.backward()
    q.put({"rank2": rank, "df2": f2.grad.clone(), "loss_r": loss_r.detach().clone()})
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
    got = [q.get(timeout=120) for _ in range(9)]
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
    for x in got:
        if "rank3" in x:
            torch.testing.assert_close(x["sparse"], x["dense"])
            torch.testing.assert_close(x["other"], torch.full((3,), 3.0))
        if "rank3b" in x:
            torch.testing.assert_close(x["sparse"], x["dense"])
            assert x["gathered"] == sorted([2, 2, 9, 7, 3, 18, 18, 9])
    # local-anchor partition: mean over ranks of the per-rank losses == the global loss, and the reduce-scatter hands
    # every rank the same world x gradient of its slice as the redundant formulation above
    sh = {x["rank2"]: x for x in got if "rank2" in x}
    torch.testing.assert_close(0.5 * (sh[0]["loss_r"] + sh[1]["loss_r"]), ref.detach(), rtol=1e-5, atol=1e-6)
    for r in range(2):
        torch.testing.assert_close(sh[r]["df2"] / 2, F.grad[r * 3:(r + 1) * 3], rtol=1e-5, atol=1e-6)
