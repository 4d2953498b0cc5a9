cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_parallel_nccl_gpu.py -x -q > gpurun_out/r2_nccl_test3.log 2>&1; tail -3 gpurun_out/r2_nccl_test3.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --quick --pad-steps 20"
for c in 8 16 32; do timeout 400 $RUN --multimem 1 --multimem-ctas $c > gpurun_out/dp2_mm_c$c.json 2> gpurun_out/dp2_mm_c$c.err; echo rc=$?; done
timeout 400 $RUN --multimem 1 --multimem-ctas 16 --sink-group 1 > gpurun_out/dp2_mm_c16_sg1.json 2>/dev/null
tail -q -n 1 gpurun_out/dp2_mm_c8.json gpurun_out/dp2_mm_c16.json gpurun_out/dp2_mm_c32.json gpurun_out/dp2_mm_c16_sg1.json
