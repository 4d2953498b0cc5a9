# tensor-pipe utilisation of the tcgen05 kernels: run on the GPU box
cd $GRAFT_REPO_ROOT
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.avg.pct_of_peak_sustained_elapsed,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.avg.pct_of_peak_sustained_elapsed,sm__sass_inst_executed_op_utcmma.sum,sm__inst_executed_pipe_tensor.sum,sm__mem_tensor_writes_op_utcmma.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
python tools/ncu_tensor_probe.py > gpurun_out/ncu_tensor_plain.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:'gemm_tc_kernel|attn_tc' --csv --log-file gpurun_out/ncu_tensor.csv python tools/ncu_tensor_probe.py > gpurun_out/ncu_tensor.log 2>&1
echo rc=$?
