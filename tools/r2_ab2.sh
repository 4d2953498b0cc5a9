cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_rf_encoder_gpu.py -x -q -k "attention_block" > gpurun_out/r2_ab_test.log 2>&1; tail -12 gpurun_out/r2_ab_test.log
timeout 200 python tools/timeline.py --out gpurun_out/timeline_ab.csv > gpurun_out/timeline_ab.txt 2>&1; head -14 gpurun_out/timeline_ab.txt
