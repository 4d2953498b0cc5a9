cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "projector or cast_pad or embed or deferred" > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/h_pytest.log
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
run() { n=$1; shift; timeout 200 python bench.py $Q "$@" > gpurun_out/h_$n.json 2>gpurun_out/h_$n.err; echo "$n rc=$? $(tail -n1 gpurun_out/h_$n.json | cut -c1-100)"; }
run base
MMVQA_VISTOK_PG=0 run nopg
run hot --hot-only
timeout 300 python tools/timeline.py --out gpurun_out/timeline_h.csv > gpurun_out/timeline_h.txt 2>&1; echo "timeline rc=$?"
