cd $GRAFT_REPO_ROOT
python tools/ncu_mid_gemm.py > gpurun_out/ncu_mid_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/mid_gemm python tools/ncu_mid_gemm.py > gpurun_out/ncu_mid.log 2>&1
echo rc=$?
