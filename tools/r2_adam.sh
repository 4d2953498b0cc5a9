cd $GRAFT_REPO_ROOT
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
for c in 0 148 296 592; do MMVQA_ADAM_EARLY_CTAS=$c timeout 300 python bench.py $Q > gpurun_out/k_adamu_$c.json 2>/dev/null; done
timeout 300 python bench.py $Q --overlap-adam 0 > gpurun_out/k_adamu_noov.json 2>/dev/null
tail -q -n 1 gpurun_out/k_adamu_*.json
timeout 300 python -m pytest tests/test_optim_graph_gpu.py -x -q 2>&1 | tail -2
timeout 200 python tools/timeline.py --out gpurun_out/timeline_r2.csv > gpurun_out/timeline_r2.txt 2>&1; head -30 gpurun_out/timeline_r2.txt
