set -x
cd $GRAFT_REPO_ROOT
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
python bench.py $Q > gpurun_out/k_base.json 2>gpurun_out/k_base.err
MMVQA_ADAM_EARLY_CTAS=32 python bench.py $Q > gpurun_out/k_adam32.json 2>/dev/null
MMVQA_ADAM_EARLY_CTAS=74 python bench.py $Q > gpurun_out/k_adam74.json 2>/dev/null
MMVQA_ADAM_EARLY_CTAS=148 python bench.py $Q > gpurun_out/k_adam148.json 2>/dev/null
python bench.py $Q --sink-group-1gpu 3 > gpurun_out/k_sg3.json 2>/dev/null
python bench.py $Q --overlap-adam 0 > gpurun_out/k_noov.json 2>/dev/null
MMVQA_RF_ENCODER=1 python bench.py $Q > gpurun_out/k_rfenc.json 2>/dev/null
python bench.py $Q --dropout 0 > gpurun_out/k_nodrop.json 2>/dev/null
tail -n 3 gpurun_out/k_*.json
