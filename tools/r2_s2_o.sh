cd $GRAFT_REPO_ROOT
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
run() { n=$1; shift; timeout 200 python bench.py $Q "$@" > gpurun_out/o_$n.json 2>gpurun_out/o_$n.err; echo "$n rc=$? $(tail -n1 gpurun_out/o_$n.json | cut -c1-100)"; }
run base
MMVQA_ADAM_EARLY_CTAS=148 run adam148
MMVQA_ADAM_EARLY_CTAS=100 run adam100
timeout 600 python -m pytest tests/test_optim_graph_gpu.py tests/test_parity_golden_gpu.py -q -m gpu -x > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/o_pytest.log
