cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_optim_graph_gpu.py -q -m gpu -x > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s_pytest.log
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
run() { n=$1; shift; timeout 200 python bench.py $Q "$@" > gpurun_out/s_$n.json 2>gpurun_out/s_$n.err; echo "$n rc=$? $(tail -n1 gpurun_out/s_$n.json | cut -c1-100)"; }
run base
timeout 300 python tools/timeline.py --out gpurun_out/timeline_s.csv > gpurun_out/timeline_s.txt 2>&1; echo "timeline rc=$?"
