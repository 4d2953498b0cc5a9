cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-eager-bar --pad-steps 0 > gpurun_out/final_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 200 python tools/profile_targets.py > gpurun_out/final_targets.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'vistok_pg_kernel|gemm_tc_kernel|attn_tc_bwd|attn_tc_fwd|cast_pad_multi|ln_bwd_packed' --launch-skip 20 -c 12 -o gpurun_out/r02_full python tools/profile_targets.py > gpurun_out/final_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/timeline.py --out gpurun_out/r02_timeline.csv > gpurun_out/r02_timeline_summary.txt 2>&1; echo "timeline rc=$?"
timeout 300 python tools/kernel_bench.py > gpurun_out/r02_kernel_bench.txt 2>&1; echo "kb rc=$?"
tail -c 600 gpurun_out/final_bench.json
