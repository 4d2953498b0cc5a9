cd $GRAFT_REPO_ROOT
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
run() { n=$1; shift; timeout 200 python bench.py $Q "$@" > gpurun_out/d_$n.json 2>gpurun_out/d_$n.err; echo "$n rc=$? $(tail -n1 gpurun_out/d_$n.json | cut -c1-120)"; }
MMVQA_L2_PREFETCH=0 run off
run on16
MMVQA_L2_PREFETCH_CTAS=4 run on4
MMVQA_L2_PREFETCH_CTAS=64 run on64
MMVQA_L2_PREFETCH=0 run hot_off --hot-only
run hot_on --hot-only
