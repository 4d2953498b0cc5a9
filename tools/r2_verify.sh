cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
timeout 300 bash tools/ncu_tensor.sh
timeout 900 python tools/sweep.py > gpurun_out/r2_sweep.txt 2>&1; echo "sweep rc=$?"
tail -5 gpurun_out/r2_pytest.log
