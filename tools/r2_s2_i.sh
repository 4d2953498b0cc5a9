cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/i_pytest.log
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
run() { n=$1; shift; timeout 200 python bench.py $Q "$@" > gpurun_out/i_$n.json 2>gpurun_out/i_$n.err; echo "$n rc=$? $(tail -n1 gpurun_out/i_$n.json | cut -c1-100)"; }
run base
MMVQA_EMBED_PREZERO=0 run noprezero
run hot --hot-only
timeout 300 python tools/timeline.py --out gpurun_out/timeline_i.csv > gpurun_out/timeline_i.txt 2>&1; echo "timeline rc=$?"
