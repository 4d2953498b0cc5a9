cd $GRAFT_REPO_ROOT
timeout 600 python tools/tc_probe.py > gpurun_out/tc_probe_default.log 2>&1
MMVQA_TC_KPS=2 timeout 600 python tools/tc_probe.py > gpurun_out/tc_probe_kps2.log 2>&1
MMVQA_TC_KPS=4 timeout 600 python tools/tc_probe.py > gpurun_out/tc_probe_kps4.log 2>&1
grep -c PASS gpurun_out/tc_probe_*.log; grep -h "FAIL\|TIMEOUT\|rc=" gpurun_out/tc_probe_*.log | head
MMVQA_TC_KPS=1 timeout 300 python tools/kernel_bench.py > gpurun_out/kb_kps1.txt 2>&1
timeout 300 python tools/kernel_bench.py > gpurun_out/kb_auto.txt 2>&1
timeout 300 python bench.py --steps 40 --warmup 5 --quick --pad-steps 20 > gpurun_out/k_kps_auto.json 2>/dev/null
MMVQA_TC_KPS=1 timeout 300 python bench.py --steps 40 --warmup 5 --quick --pad-steps 20 > gpurun_out/k_kps1.json 2>/dev/null
cat gpurun_out/k_kps_auto.json gpurun_out/k_kps1.json
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest2.log 2>&1; tail -5 gpurun_out/r2_pytest2.log
