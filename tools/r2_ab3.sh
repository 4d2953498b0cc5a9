cd $GRAFT_REPO_ROOT
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
MMVQA_NO_OVERLAP=1 MMVQA_RF_ATTN_BWD=0 timeout 300 python bench.py $Q --overlap-adam 0 > gpurun_out/k_ser_ab0.json 2>/dev/null
MMVQA_NO_OVERLAP=1 MMVQA_RF_ATTN_BWD=1 timeout 300 python bench.py $Q --overlap-adam 0 > gpurun_out/k_ser_ab1.json 2>/dev/null
MMVQA_RF_ATTN_BWD=0 timeout 300 python bench.py $Q --overlap-adam 0 > gpurun_out/k_noad_ab0.json 2>/dev/null
MMVQA_RF_ATTN_BWD=1 timeout 300 python bench.py $Q --overlap-adam 0 > gpurun_out/k_noad_ab1.json 2>/dev/null
tail -q -n 1 gpurun_out/k_ser_ab0.json gpurun_out/k_ser_ab1.json gpurun_out/k_noad_ab0.json gpurun_out/k_noad_ab1.json
