#!/usr/bin/env python
"""bench.py -- VQA fine-tune samples/sec of the MMBERT fusion-encoder hot path (BASELINE.json configs[1]):
EffNetV2-M feature maps -> 5 visual tokens -> 12-layer RealFormer (8 x 96) -> heads -> ASLSingleLabel,
forward + backward + Adam, batch 16 per GPU, T = 28, synthetic VQA-Med-shaped inputs, random-init weights.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (CPU arm: the oracle port on the host cores)

Prints ONE JSON line (rank 0).  `value`: whole-job samples/s with inputs resident in HBM (CUDA-graph replay,
CUDA-event timed, max over ranks).  `e2e`: the same step through the public module API with HOST inputs, the
host->device copies and the device->host loss read inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

HIDDEN, LAYERS, T, NUM_CLASSES, HEADS = 768, 12, 28, 1552, 8
EFFNET_MAPS = [(24, 112), (48, 56), (80, 28), (176, 14), (512, 7)]          # (channels, side) in token order
# algorithmic GEMM FLOPs per sample: forward 4.631 G, train step = 3 x forward
STEP_GFLOP_PER_SAMPLE = 13.893


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def make_args():
    return types.SimpleNamespace(task="VQA", clinicalbert="", transformer_model="realformer",
                                 cnn_encoder="tf_efficientnetv2_m", num_vis=5, hidden_size=HIDDEN, use_relu=False,
                                 heads=HEADS, hidden_dropout_prob=0.1, n_layers=LAYERS, vocab_size=30522, dataset="VQA-Med")


def synth_batch(B, seed):
    """SURVEY.md section 8d synthetic inputs: post-activation-like feature maps, [CLS] 5x0 [SEP] question [SEP] pad."""
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(B, c, s, s, generator=g).abs_() for c, s in EFFNET_MAPS]
    ids = torch.zeros(B, T, dtype=torch.long)
    seg = torch.zeros(B, T, dtype=torch.long)
    mask = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):
        qlen = int(torch.randint(4, T - 8 + 1, (1,), generator=g))
        ids[b, 0], ids[b, 6] = 101, 102
        ids[b, 7:7 + qlen] = torch.randint(1000, 30522, (qlen,), generator=g)
        ids[b, 7 + qlen] = 102
        seg[b, 7:8 + qlen] = 1
        mask[b, :8 + qlen] = 1
    target = torch.randint(0, NUM_CLASSES, (B,), generator=g)
    return feats, ids, seg, mask, target


def build_model(seed=0):
    """Model(args) exactly as vqamed2019/train.py builds it (incl. the classifier[2] swap, train.py:137), with the
    two offline stand-ins: bert-base-shaped random BertEmbeddings and no backbone (inputs are feature maps)."""
    from transformers import BertConfig, BertModel
    from mmvqa_b200.models import image_encoding as IE
    from mmvqa_b200.models import mmbert as MM
    torch.manual_seed(seed)
    old = MM.AutoModel.from_pretrained
    MM.AutoModel.from_pretrained = staticmethod(lambda name, *a, **k: BertModel(BertConfig(num_hidden_layers=1)))
    IE.models_dict[5]["tf_efficientnetv2_m"][0] = lambda *a, **k: nn.Identity()
    try:
        model = MM.Model(make_args())
    finally:
        MM.AutoModel.from_pretrained = old
    model.classifier[2] = nn.Linear(HIDDEN, NUM_CLASSES)
    return model


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(model_state, B):
    from oracle import mmbert_oracle as O          # CPU baseline leg: the one place bench.py executes oracle/
    p = {k: v.detach().clone().float().requires_grad_(v.is_floating_point()) for k, v in model_state.items()}
    params = [v for v in p.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-5)
    feats, ids, seg, mask, target = synth_batch(B, 1234)

    def step():
        opt.zero_grad(set_to_none=True)
        logits = O.model_forward(feats, ids, seg, mask, p, encoder="realformer", n_layers=LAYERS, dataset="VQA-Med")
        loss = O.asl_single_label(logits, target)
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step


def run_cpu(model_state, B, steps, warmup, budget_s=None):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_step_fn(model_state, B)
    for _ in range(warmup):
        step()
    t0, n = time.perf_counter(), 0
    while n < steps:
        step()
        n += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return n * B / dt, dt / n * 1e3, n


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model = build_model()
    state = {k: v for k, v in model.state_dict().items()}
    sps, ms, n = run_cpu(state, a.batch, a.steps, min(a.warmup, 1), budget_s=150.0)
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "vqa_finetune_samples_per_sec", "value": sps, "unit": "samples/s", "n_gpus": a.gpus,
            "steps": n, "warmup": min(a.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(a.batch, a.gpus),
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{n} full steps of batch {a.batch} (torch fp32 oracle port of models/*.py, {cores} threads)"},
            "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


FEAT_DTYPE = "fp32"


def config_dict(B, n):
    return {"workload": "configs[1]: MMBERT EfficientNetV2-M feature maps + RealFormer-12 (8 heads x 96, hidden 768) + "
                        "ASLSingleLabel fine-tune step (fwd+bwd+Adam), T=28, num_vis=5, 1552 answer classes",
            "batch_per_gpu": B, "global_batch": B * n, "seq_len": T, "parallelism": f"dp{n}",
            "scope": "hot path: feature maps -> loss -> grads -> Adam (CNN backbone is library code, out of scope)",
            "feature_maps": FEAT_DTYPE,
            "l2": "4 rotating input batches (147 MB) and 1.3 GB of weights + Adam state are streamed every step (> 126 MB L2)"}


def gpu_eager_bar(model_state, B, steps=10):
    """The GPU-side bar of SURVEY.md section 8d: the reference arithmetic as plain eager PyTorch on the same B200 (cuBLAS /
    cuDNN / ATen kernels, ~30 launches per layer), fp32 and bf16 autocast, eager and CUDA-graphed, same batch, same
    torch.optim.Adam.  /root/reference does not exist on the GPU box, so this runs the oracle port (oracle/mmbert_oracle.py:
    the line-by-line torch restatement of models/*.py that tests/test_oracle_golden.py pins bit-exactly to vectors
    generated from the unmodified reference) -- a reported bar, never the product path."""
    from oracle import mmbert_oracle as O
    out = {}
    feats, ids, seg, mask, target = [t.cuda() if torch.is_tensor(t) else [u.cuda() for u in t] for t in synth_batch(B, 1234)]
    for name, autocast, graphed in (("fp32_eager", False, False), ("bf16_autocast_eager", True, False),
                                    ("fp32_cuda_graph", False, True), ("bf16_autocast_cuda_graph", True, True)):
        try:
            p = {k: v.detach().clone().float().cuda().requires_grad_(v.is_floating_point()) for k, v in model_state.items()}
            params = [v for v in p.values() if v.requires_grad]
            opt = torch.optim.Adam(params, lr=1e-5, capturable=graphed)

            def step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    logits = O.model_forward(feats, ids, seg, mask, p, encoder="realformer", n_layers=LAYERS, dataset="VQA-Med")
                    loss = O.asl_single_label(logits.float(), target)
                loss.backward()
                opt.step()
                return loss
            run = step
            if graphed:
                s_ = torch.cuda.Stream()
                s_.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s_):
                    for _ in range(3):
                        step()
                torch.cuda.current_stream().wait_stream(s_)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    step()
                run = g.replay
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                run()
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"samples_per_s": B / (ms * 1e-3), "ms_per_step": ms}
            del p, params, opt
        except Exception as ex:      # e.g. an op that cannot be captured: report, do not fail the bench
            out[name] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    out["what"] = ("oracle port of models/*.py as eager PyTorch on this GPU (ATen/cuBLAS), same batch, torch.optim.Adam; "
                   "dropout off (the port has none)")
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def time_kernel(fn, iters=20):
    """Average device time of one fn() launch in ms.  [L2 flush, fn] x iters is captured in a CUDA graph and timed
    with CUDA events on the launching stream; the same graph without fn is subtracted, so neither host launch
    latency nor the flush is in the number and every launch starts with a cold L2 (256 MB memset) as inside the step."""
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()

    def capture(with_fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
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
    g1, g0 = capture(True), capture(False)
    t = (run(g1) - run(g0)) / iters
    del g1, g0
    return max(t, 1e-6)


def main_gpu(a):
    import torch.distributed as dist
    import mmvqa_b200
    from mmvqa_b200 import ops
    from mmvqa_b200.graph import GraphedTrainStep
    from mmvqa_b200.models.asl_singlelabel import ASLSingleLabel
    from mmvqa_b200.optim import FusedAdam
    from mmvqa_b200.parallel import GradBuckets, broadcast_parameters

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)

    def log(msg):
        if a.verbose:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)
    if world > 1:
        if a.nccl_channels > 0:       # the exchange shares the SMs with the backward pass: fewer, fatter channels
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(a.nccl_channels))
        opts = None
        if a.nccl_high_priority:      # the exchange is on the critical path of the per-layer optimizer chain
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    mmvqa_b200.set_compute_dtype(dt)
    B = a.batch
    model = build_model().cuda().train()
    if world > 1:
        broadcast_parameters(model)
    if not a.dropout:
        for m in model.modules():
            if isinstance(m, nn.Dropout):
                m.p = 0.0
    params = [p for p in model.parameters() if p.requires_grad]
    nparams = sum(p.numel() for p in params)
    # Adam of each encoder layer runs on its own stream under the rest of the backward pass.  Data parallel
    # (--dp-mode overlapped): the layer's gradient bucket is all-reduced on that stream first, so the NVLink exchange
    # overlaps the backward too and the whole step -- NCCL kernels included -- is ONE CUDA graph.
    # --dp-mode twograph: forward/backward graph, eager bucketed all-reduce, optimizer graph (no overlap).
    from mmvqa_b200.parallel import LayerwiseReducer
    bucket_dt = torch.bfloat16 if (dt == torch.bfloat16 and a.bf16_buckets) else torch.float32
    overlapped_dp = world > 1 and a.dp_mode == "overlapped"
    # in-switch all-reduce of the library (csrc/comm.cu) from 4 GPUs on: measured 3.14 ms/step against 3.37 ms for NCCL's
    # ring kernels at 8 GPUs (profiles/r02_scaling_notes.txt); at 2 GPUs a ring moves the same bytes and NCCL is 2 % faster
    use_mm = (a.multimem == 1) or (a.multimem < 0 and world >= 4)
    reducer = LayerwiseReducer(bucket_dt, multimem=use_mm, multimem_ctas=a.multimem_ctas) if overlapped_dp else None
    opt = FusedAdam(params, lr=1e-5, overlap_backward=((world == 1 or overlapped_dp) and bool(a.overlap_adam)),
                    reduce_fn=reducer, sink_group=(a.sink_group if overlapped_dp else a.sink_group_1gpu),
                    early_groups=[list(model.transformer.bert_embedding.parameters()),
                                  list(model.fc1.parameters()) + list(model.classifier.parameters())])
    crit = ASLSingleLabel()
    buckets = GradBuckets(params, dtype=bucket_dt) if (world > 1 and not overlapped_dp) else None
    if buckets is not None:
        opt.grad_scale = buckets.grad_scale
    if reducer is not None:
        opt.grad_scale = reducer.grad_scale

    step_ids = {}

    def loss_fn(f0, f1, f2, f3, f4, ids, seg, mask, target):
        step_ids["ids"] = ids                      # the graph's static input_ids tensor (row-sparse embedding exchange)
        logits, _, _ = model.forward_features([f0, f1, f2, f3, f4], ids, seg, mask)
        return crit(logits, target)
    word_w = model.transformer.bert_embedding.word_embeddings.weight
    if reducer is not None and a.sparse_embed:
        # the word-embedding gradient has at most B*T non-zero rows: exchange those instead of the dense 47 MB table
        reducer.register_row_sparse(word_w, lambda: step_ids["ids"])
    if a.sparse_embed and (world == 1 or reducer is not None):
        # ... and Adam skips the table rows that have never received gradient (their update is exactly the identity)
        opt.register_row_sparse(word_w, (lambda: step_ids["ids"]) if reducer is None else (lambda: reducer.gathered_ids(word_w)))

    NB = 4
    host = []
    for i in range(NB):
        feats, ids, seg, mask, target = synth_batch(B, 1000 * rank + i)
        if a.feat_dtype == "bf16":                 # backbone hand-off in bf16: no cast launches
            feats = [f.to(torch.bfloat16) for f in feats]
        host.append([t.pin_memory() for t in (*feats, ids, seg, mask, target)])
    dev = [[t.cuda() for t in hb] for hb in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])      # payload; the packed buffer adds < 2 KB of padding

    def step_kwargs():
        plist = [p for p in params if p.grad is not None]
        return dict(grads=buckets.grads(plist))
    log("building the graphed step")

    class _HotOnly:
        """--hot-only (tuning): zero_grad only, the captured step is forward + loss + backward"""
        def zero_grad(self, set_to_none=True):
            for p_ in params:
                p_.grad = None

        def step(self):
            pass
    if a.hot_only:
        opt.close()
    gs = GraphedTrainStep(loss_fn, dev[0], (_HotOnly() if a.hot_only else opt), warmup=3, post_backward=(buckets.pack if buckets else None),
                          eager_between=(buckets.allreduce if buckets else None),
                          step_kwargs=(step_kwargs if buckets else None),
                          capture_error_mode=("thread_local" if world > 1 else "global"),
                          main_priority=a.main_priority)
    log("graph captured: %d launches per step" % gs.launches_per_step)
    launches = gs.launches_per_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # every batch in the graph's packed input layout: one copy per step refreshes all nine input tensors
    dev_flat = [gs.pack_like(db, device=db[0].device)[0] for db in dev]
    host_flat = [gs.pack_like(hb, pin_memory=True)[0] for hb in host]
    log("timing value")
    # ---- value: inputs resident in HBM ----
    for i in range(max(a.warmup, 3)):
        gs.replay_packed(dev_flat[i % NB])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        # pad the sampled window so nvidia-smi sees the load (untimed replays; the SAME count on every rank, the
        # step contains a collective), then the timed region
        for _ in range(a.pad_steps):
            gs.replay_packed(dev_flat[0])
        barrier()
        e0.record()
        for i in range(a.steps):
            gs.replay_packed(dev_flat[i % NB])
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        for _ in range(a.pad_steps // 2):
            gs.replay_packed(dev_flat[0])
        torch.cuda.synchronize()
    loss_val = float(gs.static_loss)
    if world > 1:
        tmax = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax)
    ms_step = ms_total / a.steps
    value = B * world / (ms_step * 1e-3)

    log("timing e2e")
    # ---- e2e: host inputs, H2D + D2H inside the timed region, every step ----
    # The H2D copy of step i+1's inputs (36.7 MB of feature maps) runs on a copy stream into a staging buffer while
    # step i computes; step i+1 starts with a device-side copy staging -> static graph inputs.  Every copy and the
    # D2H loss read of every step are inside the timed region.
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    staging = torch.empty_like(gs.static_flat)
    main = torch.cuda.current_stream()
    staged, consumed = torch.cuda.Event(), torch.cuda.Event()

    def prefetch(i):
        copy_stream.wait_event(consumed)                  # staging buffer free again
        with torch.cuda.stream(copy_stream):
            staging.copy_(host_flat[i % NB], non_blocking=True)      # ONE pinned-host -> device copy per step
            staged.record(copy_stream)

    def e2e_steps(n):
        consumed.record(main)
        prefetch(0)
        for i in range(n):
            main.wait_event(staged)
            gs.static_flat.copy_(staging, non_blocking=True)
            consumed.record(main)
            if i + 1 < n:
                prefetch(i + 1)
            gs.replay_packed(gs.static_flat)
            loss_host.copy_(gs.static_loss, non_blocking=True)
            main.synchronize()                            # the loss is read on the host every step
    e2e_steps(3)
    barrier()
    e0.record()
    e2e_steps(a.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        tmax = torch.tensor([ms_e2e], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_e2e = float(tmax)
    e2e_value = B * world * a.steps / (ms_e2e * 1e-3)

    log("timed regions done")
    def finish():
        """leave the job: a CUDA graph that holds captured NCCL kernels must be gone before the communicator is
        torn down (destroy_process_group otherwise waits forever on the graph's work handles)."""
        nonlocal gs
        if world > 1:
            dist.barrier()
            if gs is not None:
                gs.close()               # drops the graph (and the NCCL work it captured) after the device has drained
                gs = None
            torch.cuda.synchronize()
            th = threading.Thread(target=dist.destroy_process_group, daemon=True)
            th.start()
            th.join(timeout=30)
            if th.is_alive():            # communicator teardown stuck on captured work: leave without it
                sys.stdout.flush()
                sys.stderr.flush()
                os._exit(0)
    if rank != 0:
        finish()
        return
    if a.quick:          # tuning runs: the two timed regions only
        print(json.dumps({"quick": True, "value": value, "ms_per_step": ms_step, "e2e": e2e_value, "ms_e2e": ms_e2e / a.steps,
                          "gpu_launches_per_step": launches, "n_gpus": world, "clocks": clocks.summary()}), flush=True)
        finish()
        return
    hbm, tf_burst, tf_sus, which = peaks()
    log("roofline measurements")
    # ---- roofline of the dominant kernel, measured live with CUDA events on the launching stream ----
    # (1) the tcgen05 GEMM family (gemm_tc_kernel / vistok_kernel): one eager forward+backward records every
    #     mmvqa_gemm launch of the step; each DISTINCT problem is then replayed alone (graph-captured back-to-back
    #     launches, cold L2, CUDA events) and the per-step total is sum(count x time).
    from mmvqa_b200 import ops as _ops
    opt.close()                    # rank 0 works alone from here on: no gradient sink, no collective
    opt.reduce_fn = None
    opt.zero_grad(set_to_none=True)
    # (captured BEFORE any eager backward on the default stream: AccumulateGrad nodes born there would tie the legacy
    # stream to the capture)
    # ---- hot path only: feature maps -> loss -> all gradients, no optimizer ----
    log("hot path only")
    hot = None
    try:
        class _NoOptimizer:
            """zero_grad only: the captured step is forward + loss + backward, nothing else"""
            def zero_grad(self, set_to_none=True):
                for p_ in params:
                    p_.grad = None

            def step(self):
                pass
        gh = GraphedTrainStep(loss_fn, dev[0], _NoOptimizer(), warmup=2, main_priority=a.main_priority)
        for i in range(5):
            gh.replay_packed(dev_flat[i % NB])
        torch.cuda.synchronize()
        e0.record()
        for i in range(a.steps):
            gh.replay_packed(dev_flat[i % NB])
        e1.record()
        e1.synchronize()
        ms_hot = e0.elapsed_time(e1) / a.steps
        hot = {"ms_per_step": ms_hot, "samples_per_s": B / (ms_hot * 1e-3), "launches_per_step": gh.launches_per_step,
               "tensor_frac": B / (ms_hot * 1e-3) * STEP_GFLOP_PER_SAMPLE * 1e9 / (tf_burst * 1e12),
               "what": "projector -> encoder -> heads -> ASL, forward + backward (every gradient), CUDA-graph replay, no Adam"}
        gh.close()
        del gh
    except Exception as ex:
        import traceback
        traceback.print_exc(file=sys.stderr)
        hot = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
    opt.zero_grad(set_to_none=True)
    _ops.gemm_record(True)
    loss_fn(*dev[0]).backward()
    rec = _ops.gemm_record(False)
    torch.cuda.synchronize()
    uniq = {}
    for sig, fl, args, keep in rec:
        ent = uniq.setdefault(sig, [0, fl, args, keep])
        ent[0] += 1
    gemm_ms, g_flops, gemm_n = 0.0, 0.0, 0
    for sig, (cnt, fl, args, keep) in uniq.items():
        t = time_kernel(lambda: _ops.gemm_replay(args), iters=10)
        gemm_ms += cnt * t
        g_flops += cnt * fl
        gemm_n += cnt
    gemm_tflops = g_flops / (gemm_ms * 1e-3) / 1e12
    M = B * T
    x = torch.randn(M, HIDDEN, device="cuda").to(dt)
    w = torch.randn(4 * HIDDEN, HIDDEN, device="cuda").to(dt)
    bias = torch.randn(4 * HIDDEN, device="cuda")
    y = torch.empty(M, 4 * HIDDEN, device="cuda", dtype=dt)
    pre = torch.empty_like(y)
    from mmvqa_b200._lib import ACT_SERF, EPI_ACT
    ms_ff1 = time_kernel(lambda: ops.gemm(M, 4 * HIDDEN, HIDDEN, x, HIDDEN, False, w, HIDDEN, False, y, 4 * HIDDEN, bias=bias,
                                          epilogue=EPI_ACT, act=ACT_SERF, aux_out=pre, ld_aux_out=4 * HIDDEN))
    ff1_tflops = 2.0 * M * 4 * HIDDEN * HIDDEN / (ms_ff1 * 1e-3) / 1e12
    xb_, wb_ = torch.randn(8192, 8192, device="cuda").to(dt), torch.randn(8192, 8192, device="cuda").to(dt)
    yb_ = torch.empty(8192, 8192, device="cuda", dtype=dt)
    ms_big = time_kernel(lambda: ops.gemm(8192, 8192, 8192, xb_, 8192, False, wb_, 8192, False, yb_, 8192), iters=5)
    big_tflops = 2.0 * 8192 ** 3 / (ms_big * 1e-3) / 1e12
    del xb_, wb_, yb_
    # (2) Adam: 16 B read (p, m, v, g) + 12 B written (p, m, v) per parameter (+2 B bf16 operand copy where cached)
    # (timed alone over ALL parameters; inside the step the per-layer updates run underneath the backward pass)
    plist_all = [p for p in params if p.grad is not None]
    opt._row_gate = {}             # the roofline run updates EVERY row: 28 B for each of the 91 M parameters
    opt._tables.clear()
    from mmvqa_b200.optim import CHUNK_BACKGROUND
    table, nchunks = opt._table("roofline", plist_all, [p.grad for p in plist_all], CHUNK_BACKGROUND)   # the bulk mode the step uses
    nparams_adam = sum(p.numel() for p in plist_all)
    grp = opt.param_groups[0]
    ms_adam = time_kernel(lambda: ops.adam_step(table, nchunks, grp["lr"], grp["betas"][0], grp["betas"][1], grp["eps"],
                                                grp["weight_decay"], 0, opt._step_dev, opt.grad_scale, 0, True), iters=10)
    adam_gbs = nparams_adam * 28 / (ms_adam * 1e-3) / 1e9
    # DRAM bytes per launch of the same kernel family from the committed ncu launch list (profiles/, not measured here)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roof = {"kernel": "gemm_tc_kernel (tcgen05/TMEM/TMA bf16 GEMM family: %d launches per step, %.0f%% of the step time)"
                      % (gemm_n, 100.0 * gemm_ms / ms_step),
            "bound": "tensor", "achieved": gemm_tflops, "peak": tf_burst, "unit": "TFLOP/s", "frac": gemm_tflops / tf_burst,
            "traffic": traffic, "traffic_source": "profiles/%s (ncu dram__bytes_read+write per launch, family average)" % os.path.basename(tpath).replace("gemm_traffic.json", "launches_summary.txt"),
            "peak_source": which + " (MEASURED_PEAKS.json bf16_tflops, burst: kernels timed alone)",
            "flops_per_launch": g_flops / max(gemm_n, 1), "ms_per_launch": gemm_ms / max(gemm_n, 1),
            "how": "every mmvqa_gemm problem of one forward+backward at the bench shape (%d distinct), each replayed alone "
                   "and timed with CUDA events (cold L2); total = sum(count x time); algorithmic FLOPs = 2MNK" % len(uniq),
            "same_kernel_other_shapes": {
                "ff1_448x3072x768_bias_serf": {"achieved": ff1_tflops, "frac": ff1_tflops / tf_burst, "ms_per_launch": ms_ff1},
                "square_8192": {"achieved": big_tflops, "frac": big_tflops / tf_burst, "ms_per_launch": ms_big}},
            "hbm_kernel": {"kernel": "adam_kernel (multi-tensor Adam, %.1f M fp32 params, 28 B/param)" % (nparams / 1e6),
                           "bound": "hbm", "achieved": adam_gbs, "peak": hbm, "unit": "GB/s", "frac": adam_gbs / hbm,
                           "ms_per_launch": ms_adam, "share_of_step": ms_adam / ms_step},
            "step_tensor_frac": value / world * STEP_GFLOP_PER_SAMPLE * 1e9 / (tf_burst * 1e12)}
    eager = None
    if world == 1 and not a.no_eager_bar:
        log("eager PyTorch bar")
        eager = gpu_eager_bar({k: v.detach() for k, v in model.state_dict().items()}, B)
    cpu = None
    if world == 1 and not a.no_cpu:
        state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        sps, ms_cpu, n = run_cpu(state, B, 1000, 1, budget_s=15.0)
        cores = os.cpu_count() or 1
        cpu = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "ms_per_step": ms_cpu,
               "sample": f"{n} full steps of the same batch-{B} workload (torch fp32 oracle port, {cores} threads)"}
    line = {"metric": "vqa_finetune_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic", "config": config_dict(B, world),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": launches * a.steps, "gpu_launches_per_step": launches, "clocks": clocks.summary(),
            "roofline": roof, "cpu_baseline": cpu, "hot_path_only": hot, "gpu_eager_reference": eager,
            "loss": loss_val, "dropout": bool(a.dropout), "params": nparams}
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=16, help="samples per GPU (weak scaling)")
    ap.add_argument("--dropout", type=int, default=1, help="1: training-mode dropout on (throughput runs), 0: off")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eager-bar", action="store_true", help="skip the eager-PyTorch-on-GPU bar")
    ap.add_argument("--verbose", action="store_true", help="progress lines on stderr")
    ap.add_argument("--bf16-buckets", type=int, default=1, help="data parallel: all-reduce gradients as bf16 (bf16 path only)")
    ap.add_argument("--overlap-adam", type=int, default=1, help="1 GPU: update each layer under the rest of the backward pass")
    ap.add_argument("--nccl-channels", type=int, default=0, help="cap NCCL channels (CTAs) per collective; 0 = NCCL default")
    ap.add_argument("--main-priority", type=int, default=-1, help="CUDA priority of the captured main stream (< 0 = above the side / optimizer streams)")
    ap.add_argument("--nccl-high-priority", type=int, default=0)
    ap.add_argument("--sink-group", type=lambda v: [int(x) for x in v.split(",")], default=[4],
                    help="data parallel: encoder layers per all-reduce + Adam launch; a comma list is a schedule (last entry repeats)")
    ap.add_argument("--sink-group-1gpu", type=int, default=1, help="single GPU: encoder layers per Adam launch")
    ap.add_argument("--dp-mode", default="overlapped", choices=["overlapped", "twograph"])
    ap.add_argument("--multimem", type=int, default=-1, help="data parallel: in-switch all-reduce kernel of the library (symmetric memory) instead of NCCL; -1 = from 4 GPUs on")
    ap.add_argument("--multimem-ctas", type=int, default=16)
    ap.add_argument("--sparse-embed", type=int, default=1, help="exchange (data parallel) and update (Adam row gate) only the touched word-embedding rows; exact")
    ap.add_argument("--feat-dtype", default="fp32", choices=["fp32", "bf16"], help="dtype of the feature maps handed to the path (fp32 = as the reference's backbone emits them)")
    ap.add_argument("--quick", action="store_true", help="tuning: print value / e2e only (no roofline, no CPU leg)")
    ap.add_argument("--hot-only", action="store_true", help="tuning (with --quick): no optimizer in the captured step")
    ap.add_argument("--pad-steps", type=int, default=100, help="untimed steps around the timed region (clock sampling)")
    a = ap.parse_args()
    if os.environ.get("MMVQA_BENCH_WATCHDOG"):      # debugging aid: dump every thread's stack and exit if the run hangs
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["MMVQA_BENCH_WATCHDOG"]), exit=True)
    FEAT_DTYPE = a.feat_dtype
    if a.impl == "reference":
        main_reference(a)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        main_gpu(a)
