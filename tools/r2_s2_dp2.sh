cd $GRAFT_REPO_ROOT
export MMVQA_BENCH_WATCHDOG=240
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --quick --pad-steps 20 > gpurun_out/dp2_quick.json 2> gpurun_out/dp2_quick.err; echo "dp2 rc=$?"; tail -n1 gpurun_out/dp2_quick.json | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 5 --quick --pad-steps 20 --multimem 1 > gpurun_out/dp2_mm.json 2> gpurun_out/dp2_mm.err; echo "dp2mm rc=$?"; tail -n1 gpurun_out/dp2_mm.json | cut -c1-200
timeout 400 python -m pytest tests/test_parallel_nccl_gpu.py -q -m gpu -x > gpurun_out/dp2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/dp2_pytest.log
