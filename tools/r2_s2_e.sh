cd $GRAFT_REPO_ROOT
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
run() { n=$1; shift; timeout 200 python bench.py $Q "$@" > gpurun_out/e_$n.json 2>gpurun_out/e_$n.err; echo "$n rc=$? $(tail -n1 gpurun_out/e_$n.json | cut -c1-100)"; }
run base
MMVQA_ATTN_BWD_WARPS=2 run aw2
MMVQA_WGRAD_SOLO=1 run solo1
MMVQA_WGRAD_SOLO=3 run solo3
MMVQA_WGRAD_SOLO=1 MMVQA_ATTN_BWD_WARPS=2 run solo1_aw2
run hot_base --hot-only
MMVQA_ATTN_BWD_WARPS=2 run hot_aw2 --hot-only
MMVQA_WGRAD_SOLO=1 run hot_solo1 --hot-only
MMVQA_WGRAD_SOLO=1 MMVQA_ATTN_BWD_WARPS=2 run hot_solo1_aw2 --hot-only
