cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_rf_encoder_gpu.py -x -q -k "attention_block" > gpurun_out/r2_ab_test.log 2>&1; tail -25 gpurun_out/r2_ab_test.log
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
MMVQA_RF_ATTN_BWD=0 timeout 300 python bench.py $Q > gpurun_out/k_ab0.json 2>/dev/null
MMVQA_RF_ATTN_BWD=1 timeout 300 python bench.py $Q > gpurun_out/k_ab1.json 2>gpurun_out/k_ab1.err
tail -q -n 1 gpurun_out/k_ab0.json gpurun_out/k_ab1.json; tail -3 gpurun_out/k_ab1.err
