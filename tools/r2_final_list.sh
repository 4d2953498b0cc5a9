cd $GRAFT_REPO_ROOT
timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu --no-eager-bar --pad-steps 0 > gpurun_out/list_plain.log 2>&1; echo "plain rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-eager-bar --pad-steps 0 > gpurun_out/final_ncu_list.log 2>&1; echo "ncu list rc=$?"
