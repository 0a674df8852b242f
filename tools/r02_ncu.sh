#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2_kernel|wgrad_tc_kernel" -s 3 -c 9 -f -o gpurun_out/r02_ncu_full_tc python tools/ncu_harness.py > gpurun_out/r02_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -s 1400 -c 430 --csv --log-file gpurun_out/r02_step_metrics_ncu.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r02_ncu_list.log 2>&1
python tools/kernel_table.py gpurun_out/r02_step_metrics_ncu.csv > gpurun_out/r02_step_metrics_ncu_summary.md; head -30 gpurun_out/r02_step_metrics_ncu_summary.md; ls -la gpurun_out/r02_ncu_full_tc.ncu-rep
