set -x
timeout 120 python tools/one_step.py > gpurun_out/r2_one_step_plain.log 2>&1 || exit 1
timeout 900 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --csv --log-file gpurun_out/r2_launches.csv python tools/one_step.py > gpurun_out/r2_ncu_list.log 2>&1; echo rc=$? >> gpurun_out/r2_ncu_list.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_tc2_kernel<64, 256, 4" -c 3 -f -o gpurun_out/r2_conv256 python tools/one_step.py > gpurun_out/r2_ncu_conv.log 2>&1; echo rc=$? >> gpurun_out/r2_ncu_conv.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"nbf_apply_kernel|nbf_bwd3_kernel|nb_cl_fwd_kernel" -c 6 -f -o gpurun_out/r2_nb python tools/one_step.py > gpurun_out/r2_ncu_nb.log 2>&1; echo rc=$? >> gpurun_out/r2_ncu_nb.log
ls -la gpurun_out/*.ncu-rep
