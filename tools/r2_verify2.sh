cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log; tail -4 gpurun_out/r2_pytest3.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 200 python tools/timeline.py --out gpurun_out/timeline_r2b.csv > gpurun_out/timeline_r2b.txt 2>&1
