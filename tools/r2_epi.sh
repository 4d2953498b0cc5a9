cd $GRAFT_REPO_ROOT
for v in 0; do echo "== default"; timeout 300 python tools/gemm_trace.py --big 2>&1 | grep -E "^(mid|big|wgrad|ff|proj|kqv)|span|per-CTA|epilogue detail"; done > gpurun_out/epi_probe.txt
echo "== MMVQA_TC_BN=256" >> gpurun_out/epi_probe.txt
MMVQA_TC_BN=256 timeout 300 python tools/gemm_trace.py --big 2>&1 | grep -E "^(mid|big|wgrad|ff|proj|kqv)|span|per-CTA|epilogue detail" | grep -A2 "^mid" >> gpurun_out/epi_probe.txt
timeout 600 python tools/tc_probe.py > gpurun_out/tc_probe.txt 2>&1; tail -3 gpurun_out/tc_probe.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_golden_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --quick --no-eager-bar 2>&1 | tail -1 | cut -c1-400
