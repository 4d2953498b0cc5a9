cd $GRAFT_REPO_ROOT
for cfg in "64 2" "64 4" "128 2" "32 4" "32 2"; do set -- $cfg; MMVQA_TC_BN=$1 MMVQA_TC_KPS=$2 timeout 200 python tools/kernel_bench.py 2>/dev/null | head -16 | sed "s/^/BN$1 KPS$2 | /" ; done > gpurun_out/kb_bn_sweep.txt
timeout 200 python tools/kernel_bench.py 2>/dev/null | head -16 | sed "s/^/auto | /" >> gpurun_out/kb_bn_sweep.txt
