cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_parallel_nccl_gpu.py -x -q > gpurun_out/r2_nccl_test.log 2>&1; tail -3 gpurun_out/r2_nccl_test.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --quick --pad-steps 20"
timeout 400 $RUN > gpurun_out/dp2_sparse.json 2> gpurun_out/dp2_sparse.err; echo rc=$?
timeout 400 $RUN --sparse-embed 0 > gpurun_out/dp2_dense.json 2> gpurun_out/dp2_dense.err; echo rc=$?
timeout 400 $RUN --feat-dtype bf16 > gpurun_out/dp2_bf16maps.json 2> gpurun_out/dp2_bf16maps.err; echo rc=$?
tail -n 2 gpurun_out/dp2_*.json
