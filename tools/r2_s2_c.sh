cd $GRAFT_REPO_ROOT
Q="--steps 40 --warmup 5 --quick --pad-steps 20"
for v in "base:" "noov:--overlap-adam 0" "hot:--hot-only" ; do
  n=${v%%:*}; f=${v#*:}
  timeout 200 python bench.py $Q $f > gpurun_out/c_$n.json 2>gpurun_out/c_$n.err; echo "$n rc=$?"; tail -n1 gpurun_out/c_$n.json | cut -c1-200
done
MMVQA_NO_OVERLAP=1 timeout 200 python bench.py $Q --hot-only > gpurun_out/c_hot_noside.json 2>gpurun_out/c_hot_noside.err; tail -n1 gpurun_out/c_hot_noside.json | cut -c1-200
MMVQA_NO_OVERLAP=1 timeout 200 python bench.py $Q > gpurun_out/c_noside.json 2>gpurun_out/c_noside.err; tail -n1 gpurun_out/c_noside.json | cut -c1-200
MMVQA_ADAM_EARLY_CTAS=48 timeout 200 python bench.py $Q > gpurun_out/c_adam48.json 2>/dev/null; tail -n1 gpurun_out/c_adam48.json | cut -c1-200
