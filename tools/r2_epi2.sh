cd $GRAFT_REPO_ROOT
for k in 0 1; do for bn in 0 256; do echo "== MMVQA_TC_KPS=$k MMVQA_TC_BN=$bn"; MMVQA_TC_KPS=$k MMVQA_TC_BN=$bn timeout 300 python tools/gemm_trace.py --big 2>&1 | grep -E "^(mid|big|wgrad)|per-CTA" | grep -A1 "^mid\|^wgrad"; done; done > gpurun_out/epi_probe2.txt
