cd $GRAFT_REPO_ROOT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 40 --warmup 5 --quick --pad-steps 20"
timeout 300 $RUN > gpurun_out/dp8_nccl_sparse.json 2> gpurun_out/dp8_nccl_sparse.err; echo rc=$?
timeout 300 $RUN --sparse-embed 0 > gpurun_out/dp8_nccl_dense.json 2>/dev/null; echo rc=$?
timeout 300 $RUN --multimem 1 --multimem-ctas 32 > gpurun_out/dp8_mm32.json 2> gpurun_out/dp8_mm32.err; echo rc=$?
timeout 300 $RUN --multimem 1 --multimem-ctas 64 > gpurun_out/dp8_mm64.json 2>/dev/null; echo rc=$?
timeout 300 $RUN --multimem 1 --multimem-ctas 16 > gpurun_out/dp8_mm16.json 2>/dev/null; echo rc=$?
timeout 300 $RUN --multimem 1 --multimem-ctas 32 --feat-dtype bf16 > gpurun_out/dp8_mm32_bf16maps.json 2>/dev/null; echo rc=$?
tail -q -n 1 gpurun_out/dp8_*.json
