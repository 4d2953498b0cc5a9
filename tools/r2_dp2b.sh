cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_parallel_nccl_gpu.py -x -q > gpurun_out/r2_nccl_test2.log 2>&1; tail -15 gpurun_out/r2_nccl_test2.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --quick --pad-steps 20"
timeout 400 $RUN --multimem 1 > gpurun_out/dp2_mm.json 2> gpurun_out/dp2_mm.err; echo rc=$?
timeout 400 $RUN --multimem 1 --sink-group 1 > gpurun_out/dp2_mm_sg1.json 2> gpurun_out/dp2_mm_sg1.err; echo rc=$?
timeout 400 $RUN --multimem 0 > gpurun_out/dp2_nccl.json 2> gpurun_out/dp2_nccl.err; echo rc=$?
tail -n 1 gpurun_out/dp2_mm.json gpurun_out/dp2_mm_sg1.json gpurun_out/dp2_nccl.json; tail -5 gpurun_out/dp2_mm.err
