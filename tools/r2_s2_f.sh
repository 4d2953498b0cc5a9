cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_optim_graph_gpu.py -x -q -m gpu -k "embed or adam or graph" > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/f_pytest.log
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
run() { n=$1; shift; timeout 200 python bench.py $Q "$@" > gpurun_out/f_$n.json 2>gpurun_out/f_$n.err; echo "$n rc=$? $(tail -n1 gpurun_out/f_$n.json | cut -c1-100)"; }
run base
run nosparse --sparse-embed 0
MMVQA_WGRAD_SOLO=3 run solo3
timeout 300 python tools/timeline.py --out gpurun_out/timeline_f.csv > gpurun_out/timeline_f.txt 2>&1; echo "timeline rc=$?"
