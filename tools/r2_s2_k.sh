cd $GRAFT_REPO_ROOT
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
run() { n=$1; shift; timeout 200 python bench.py $Q "$@" > gpurun_out/k_$n.json 2>gpurun_out/k_$n.err; echo "$n rc=$? $(tail -n1 gpurun_out/k_$n.json | cut -c1-100)"; }
run late
MMVQA_WGRAD_LATE=0 run early
run hot_late --hot-only
MMVQA_WGRAD_LATE=0 run hot_early --hot-only
timeout 300 python -m pytest tests/test_rf_encoder_gpu.py tests/test_parity_golden_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/k_pytest.log
timeout 300 python tools/timeline.py --out gpurun_out/timeline_k.csv > gpurun_out/timeline_k.txt 2>&1; echo "timeline rc=$?"
