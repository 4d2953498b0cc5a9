cd $GRAFT_REPO_ROOT
timeout 600 python tools/tc_probe.py > gpurun_out/tc_probe_stage.log 2>&1; grep -c PASS gpurun_out/tc_probe_stage.log; grep -v PASS gpurun_out/tc_probe_stage.log | head -5
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_golden_gpu.py tests/test_fullsize_gpu.py -x -q > gpurun_out/r2_pytest_stage.log 2>&1; tail -3 gpurun_out/r2_pytest_stage.log
timeout 300 python tools/kernel_bench.py > gpurun_out/kb_stage.txt 2>&1
MMVQA_TC_NO_STAGE=1 timeout 300 python tools/kernel_bench.py > gpurun_out/kb_nostage.txt 2>&1
timeout 200 python tools/gemm_trace.py --big 2>&1 | sed -n '/^mid/,$p' > gpurun_out/gemm_trace_big_stage.txt
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
timeout 300 python bench.py $Q > gpurun_out/k_stage.json 2>/dev/null
MMVQA_TC_NO_STAGE=1 timeout 300 python bench.py $Q > gpurun_out/k_nostage.json 2>/dev/null
tail -q -n 1 gpurun_out/k_stage.json gpurun_out/k_nostage.json
