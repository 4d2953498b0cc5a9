cd $GRAFT_REPO_ROOT
export MMVQA_BENCH_WATCHDOG=200
timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 40 --warmup 5 --quick --pad-steps 20 > gpurun_out/dp8_quick.json 2> gpurun_out/dp8_quick.err; echo "dp8 rc=$?"; tail -n1 gpurun_out/dp8_quick.json | cut -c1-220
timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 40 --warmup 5 --quick --pad-steps 20 --feat-dtype bf16 > gpurun_out/dp8_bf16.json 2> gpurun_out/dp8_bf16.err; echo "dp8 bf16 rc=$?"; tail -n1 gpurun_out/dp8_bf16.json | cut -c1-220
