cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_optim_graph_gpu.py tests/test_rf_encoder_gpu.py tests/test_parity_golden_gpu.py -x -q -m gpu > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/g_pytest.log
python tools/adam_gate_bench.py 2>&1 | tail -2
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
run() { n=$1; shift; timeout 200 python bench.py $Q "$@" > gpurun_out/g_$n.json 2>gpurun_out/g_$n.err; echo "$n rc=$? $(tail -n1 gpurun_out/g_$n.json | cut -c1-100)"; }
run base
MMVQA_LN_DEFER=0 run nodefer
MMVQA_WGRAD_SOLO=3 run solo3
run hot --hot-only
timeout 300 python tools/timeline.py --out gpurun_out/timeline_g.csv > gpurun_out/timeline_g.txt 2>&1; echo "timeline rc=$?"
