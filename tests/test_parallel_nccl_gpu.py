"""NCCL (2 ranks, 2 GPUs) tests of the two exchanges of SURVEY.md section 8e: the SupCon feature gather (forward order,
backward reduction) and the local-anchor-rows SupCon loss with its reduce-scatter backward, against the oracle on the
global batch, plus the layer-wise gradient reducer.  Skipped on a box with fewer than two GPUs (one NCCL rank per GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mmbert_oracle as O

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import mmvqa_b200
    from mmvqa_b200.models.SupConLoss.loss import SupConLoss
    from mmvqa_b200.parallel import LayerwiseReducer, gather_mask_rows, gather_supcon_features, supcon_loss_sharded
    n = 64                                               # samples per rank, 2 views, D = 128 (C4 shape, scaled down)
    g = torch.Generator().manual_seed(7)
    F = torch.randn(world * n, 2, 128, generator=g)
    F = F / F.norm(dim=-1, keepdim=True)
    soft = torch.rand(world * n, world * n, generator=g)
    soft.fill_diagonal_(1.0)
    out = {"rank": rank}
    for dt in (torch.float32, torch.bfloat16):
        with mmvqa_b200.compute_dtype_scope(dt):
            f1 = F[rank * n:(rank + 1) * n].cuda().requires_grad_(True)
            gathered = gather_supcon_features(f1)
            full_mask = gather_mask_rows(soft[rank * n:(rank + 1) * n].cuda())
            l1 = SupConLoss()(gathered, mask=full_mask)
            l1.backward()
            f2 = F[rank * n:(rank + 1) * n].cuda().requires_grad_(True)
            l2 = supcon_loss_sharded(f2, mask=full_mask)
            l2.backward()
            l2m = l2.detach().clone()
            dist.all_reduce(l2m)
            key = "fp32" if dt == torch.float32 else "bf16"
            out[key] = {"gathered": gathered.detach().cpu(), "mask": full_mask.cpu(), "l1": l1.item(), "df1": f1.grad.cpu(),
                        "l2_mean": l2m.item() / world, "df2": f2.grad.cpu()}
    red = LayerwiseReducer(torch.bfloat16)
    gr = [torch.full((1000,), float(rank + 1), device="cuda"), torch.full((7,), 2.0 * (rank + 1), device="cuda")]
    params = [torch.nn.Parameter(torch.zeros(1000, device="cuda")), torch.nn.Parameter(torch.zeros(7, device="cuda"))]
    views = red(params, gr)
    out["reduced"] = [v.float().cpu() * red.grad_scale for v in views]
    # the library's in-switch all-reduce (multimem.ld_reduce / multimem.st over symmetric memory), bf16 and fp32 buckets,
    # called repeatedly (the signal pads must return to their idle state) and with an odd element count
    for dt, key in ((torch.bfloat16, "mm_bf16"), (torch.float32, "mm_fp32")):
        redm = LayerwiseReducer(dt, multimem=True, multimem_ctas=8)
        gen = torch.Generator(device="cuda").manual_seed(3 + rank)
        res = []
        for it in range(3):
            gm = [torch.randn(300001, device="cuda", generator=gen), torch.randn(77, device="cuda", generator=gen)]
            pm = params if it else [torch.nn.Parameter(torch.zeros(300001, device="cuda")), torch.nn.Parameter(torch.zeros(77, device="cuda"))]
            if it == 0:
                params = pm
            vm = redm(params, gm)
            torch.cuda.synchronize()
            exp = [g.to(dt).float() for g in gm]
            for e in exp:
                dist.all_reduce(e)
            res.append([(v.float() - e).abs().max().item() / e.abs().max().item() for v, e in zip(vm, exp)])
        out[key] = {"err": res, "used_multimem": len(redm._handles) > 0}
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs (one NCCL rank per GPU)")
def test_supcon_exchanges_over_nccl():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        o = q.get(timeout=300)
        got[o["rank"]] = o
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    n = 64
    g = torch.Generator().manual_seed(7)
    F = torch.randn(world * n, 2, 128, generator=g)
    F = (F / F.norm(dim=-1, keepdim=True)).requires_grad_(True)
    soft = torch.rand(world * n, world * n, generator=g)
    soft.fill_diagonal_(1.0)
    ref = O.supcon_loss(F, mask=soft)
    ref.backward()
    for key, tol, gtol in (("fp32", 1e-4, 2e-3), ("bf16", 3e-2, 0.15)):
        for r in range(world):
            o = got[r][key]
            torch.testing.assert_close(o["gathered"], F.detach())                 # rank-major gather == global batch order
            torch.testing.assert_close(o["mask"], soft)
            assert abs(o["l1"] - ref.item()) < tol * abs(ref.item())
            assert abs(o["l2_mean"] - ref.item()) < tol * abs(ref.item()), "mean of the per-rank local-anchor losses"
            want = world * F.grad[r * n:(r + 1) * n]                              # DP averaging divides by world again
            for df in (o["df1"], o["df2"]):
                e = ((df - want).abs().max() / want.abs().max()).item()
                assert e < gtol, (key, r, e)
    for r in range(world):
        for key, tol in (("mm_bf16", 1e-2), ("mm_fp32", 1e-6)):
            assert got[r][key]["used_multimem"], "B200 / NVSwitch boxes support NVLink multicast: the NCCL fallback was taken"
            for errs in got[r][key]["err"]:
                assert max(errs) < tol, (key, errs)
        torch.testing.assert_close(got[r]["reduced"][0], torch.full((1000,), 1.5))
        torch.testing.assert_close(got[r]["reduced"][1], torch.full((7,), 3.0))
