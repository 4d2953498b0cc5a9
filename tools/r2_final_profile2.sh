cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'vistok_pg_kernel|gemm_tc_kernel|attn_tc_bwd|attn_tc_fwd|cast_pad_multi|ln_bwd_packed' --launch-skip 10 -c 9 -o gpurun_out/r02_full python tools/profile_targets.py > gpurun_out/final_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
