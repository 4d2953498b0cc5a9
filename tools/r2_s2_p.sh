cd $GRAFT_REPO_ROOT
for T in 28 75 128; do KB_ATTN_ONLY=1 timeout 120 python tools/kernel_bench.py 16 $T 2>&1 | grep rf_attn; done > gpurun_out/p_attn_rows.txt
cat gpurun_out/p_attn_rows.txt
