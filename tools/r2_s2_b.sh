cd $GRAFT_REPO_ROOT
timeout 300 python tools/kernel_bench.py > gpurun_out/s2b_kernel_bench.txt 2>&1; echo "kb rc=$?"
timeout 120 python tools/ncu_chain_probe.py > gpurun_out/s2b_probe.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_tc_bwd|ln_bwd_packed|add_ln_fwd_parts|attn_tc_fwd' --launch-skip 8 -c 5 -o gpurun_out/s2b_chain python tools/ncu_chain_probe.py > gpurun_out/s2b_ncu.log 2>&1; echo "ncu rc=$?"
grep -E "attn|layernorm|add_layer" gpurun_out/s2b_kernel_bench.txt
