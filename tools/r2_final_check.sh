cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 300 python tools/timeline.py --out gpurun_out/r02_timeline.csv > gpurun_out/r02_timeline_summary.txt 2>&1; echo "timeline rc=$?"
