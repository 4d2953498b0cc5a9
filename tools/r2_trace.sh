cd $GRAFT_REPO_ROOT
timeout 200 python tools/gemm_trace.py --big 2>&1 | sed -n '/^mid/,$p' > gpurun_out/gemm_trace_big_auto.txt
MMVQA_TC_KPS=1 timeout 200 python tools/gemm_trace.py --big 2>&1 | sed -n '/^mid/,$p' > gpurun_out/gemm_trace_big_kps1.txt
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu --no-eager-bar > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench3.json').read().strip().splitlines()[-1]); print(d['hot_path_only'], d['ms_per_step'])"
