cd $GRAFT_REPO_ROOT
export MMVQA_BENCH_WATCHDOG=200
for sg in "4,4,2,1,1" "1"; do
  n=$(echo $sg | tr ',' '_')
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953${#sg} bench.py --gpus 8 --steps 40 --warmup 5 --quick --pad-steps 20 --sink-group $sg > gpurun_out/dp8_sg$n.json 2> gpurun_out/dp8_sg$n.err; echo "sg $sg rc=$?"; tail -n1 gpurun_out/dp8_sg$n.json | cut -c1-160
done
