cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_rf_encoder_gpu.py tests/test_parity_golden_gpu.py -q -m gpu -x > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/m_pytest.log
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
run() { n=$1; shift; timeout 200 python bench.py $Q "$@" > gpurun_out/m_$n.json 2>gpurun_out/m_$n.err; echo "$n rc=$? $(tail -n1 gpurun_out/m_$n.json | cut -c1-100)"; }
run tab
MMVQA_TC_NO_SERF_TAB=1 run notab
run hot_tab --hot-only
MMVQA_TC_NO_SERF_TAB=1 run hot_notab --hot-only
KB_ONLY_SPLIT=1 timeout 200 python tools/kernel_bench.py 2>&1 | head -12 > gpurun_out/m_kb_tab.txt
KB_ONLY_SPLIT=1 MMVQA_TC_NO_SERF_TAB=1 timeout 200 python tools/kernel_bench.py 2>&1 | head -12 > gpurun_out/m_kb_notab.txt
paste -d'\n' gpurun_out/m_kb_tab.txt gpurun_out/m_kb_notab.txt | grep -E "ff1 fwd|ff2 dgrad|proj fwd"
