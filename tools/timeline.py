"""Device timeline of ONE replay of the captured bench step (CUPTI kernel records through torch.profiler):
start offset, duration and stream of every kernel, plus gap / overlap statistics.  nsys is not in the image.

    python tools/timeline.py [--batch 16] [--out gpurun_out/timeline.csv]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--out", default="gpurun_out/timeline.csv")
    ap.add_argument("--dropout", type=int, default=1)
    ap.add_argument("--overlap", type=int, default=1)
    ap.add_argument("--main-priority", type=int, default=-1)
    a = ap.parse_args()
    import mmvqa_b200
    from mmvqa_b200.graph import GraphedTrainStep
    from mmvqa_b200.models.asl_singlelabel import ASLSingleLabel
    from mmvqa_b200.optim import FusedAdam
    mmvqa_b200.set_compute_dtype(torch.bfloat16)
    model = bench.build_model().cuda().train()
    if not a.dropout:
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    params = [p for p in model.parameters() if p.requires_grad]
    opt = FusedAdam(params, lr=1e-5, overlap_backward=bool(a.overlap),
                    early_groups=[list(model.transformer.bert_embedding.parameters()),
                                  list(model.fc1.parameters()) + list(model.classifier.parameters())])
    crit = ASLSingleLabel()
    step_ids = {}
    opt.register_row_sparse(model.transformer.bert_embedding.word_embeddings.weight, lambda: step_ids["ids"])

    def loss_fn(f0, f1, f2, f3, f4, ids, seg, mask, target):
        step_ids["ids"] = ids
        logits, _, _ = model.forward_features([f0, f1, f2, f3, f4], ids, seg, mask)
        return crit(logits, target)
    dev = [t.cuda() for t in (lambda b: (*b[0], *b[1:]))(bench.synth_batch(a.batch, 0))]
    gs = GraphedTrainStep(loss_fn, dev, opt, warmup=3, main_priority=a.main_priority)
    for _ in range(5):
        gs.replay(*dev)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            gs.replay(*dev)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # split into replays: each replay has the same number of kernels
    n = len(evs) // 3
    evs = evs[n:2 * n]
    t0 = evs[0].time_range.start
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    rows = []
    for e in evs:
        st, en = e.time_range.start - t0, e.time_range.end - t0
        rows.append((st, en, getattr(e, "device_resource_id", getattr(e, "device_index", 0)), e.name))
    with open(a.out, "w") as f:
        f.write("start_us,dur_us,stream,name\n")
        for st, en, sid, name in rows:
            f.write("%.2f,%.2f,%s,\"%s\"\n" % (st, en - st, sid, name[:110]))
    span = max(r[1] for r in rows)
    busy = sum(r[1] - r[0] for r in rows)
    # union of busy intervals -> idle time of the whole device
    cur_e, union = 0.0, 0.0
    for st, en, _, _ in sorted(rows):
        if en > cur_e:
            union += en - max(st, cur_e)
            cur_e = en
    print("kernels %d  span %.1f us  sum of kernel durations %.1f us  device-busy union %.1f us  idle %.1f us"
          % (len(rows), span, busy, union, span - union))
    agg = {}
    for st, en, _, name in rows:
        k = name[:70]
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += en - st
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print("%9.1f us %5.1f%% n=%4d avg %7.2f  %s" % (t, 100 * t / busy, c, t / c, k))


if __name__ == "__main__":
    main()
