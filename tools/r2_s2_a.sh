cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 40 --warmup 5 --quick --pad-steps 20 > gpurun_out/s2a_quick.json 2> gpurun_out/s2a_quick.err; echo "quick rc=$?"
timeout 300 python tools/timeline.py --out gpurun_out/timeline_s2a.csv > gpurun_out/timeline_s2a.txt 2>&1; echo "timeline rc=$?"
cat gpurun_out/s2a_quick.json; head -40 gpurun_out/timeline_s2a.txt
