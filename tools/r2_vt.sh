cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_golden_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/sweep.py --quick 2>&1 | grep projector
timeout 300 python bench.py --quick --no-eager-bar 2>&1 | tail -1 | cut -c1-300
bash tools/r2_epi2.sh
