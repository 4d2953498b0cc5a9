cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none -k regex:'vistok_pg_kernel|gemm_tc_kernel|attn_tc_bwd|attn_tc_fwd|cast_pad_multi|ln_bwd_packed' --launch-skip 10 -c 7 -o /tmp/r02_full python tools/profile_targets.py > gpurun_out/final_ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_extract.py /tmp/r02_full.ncu-rep > gpurun_out/r02_ncu_full_extract.txt 2>&1; ls -la /tmp/r02_full.ncu-rep
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
