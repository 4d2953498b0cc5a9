cd $GRAFT_REPO_ROOT
export MMVQA_BENCH_WATCHDOG=200
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 5 --quick --pad-steps 10 > gpurun_out/dp2_final.json 2> gpurun_out/dp2_final.err; echo "dp2 rc=$?"; tail -n1 gpurun_out/dp2_final.json | cut -c1-200
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 30 --warmup 5 --quick --pad-steps 10 --multimem 1 > gpurun_out/dp2_final_mm.json 2> gpurun_out/dp2_final_mm.err; echo "dp2mm rc=$?"; tail -n1 gpurun_out/dp2_final_mm.json | cut -c1-200
