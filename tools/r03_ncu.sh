#!/bin/bash
# round-3 evidence on one B200: bench line, in-graph kernel profile + launch timeline, ncu launch list, ncu --set full of the tensor-core kernels
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r03_bench_s_b64_n1.json 2> gpurun_out/r03_bench.err
python tools/graph_profile.py --timeline gpurun_out/r03_graph_timeline_step.txt > gpurun_out/r03_graph_profile_step.md 2>/dev/null
python tools/trace_gemm2.py > gpurun_out/r03_pair_gemm_timeline.log 2>&1
python tools/trace_wgrad.py > gpurun_out/r03_wgrad_timeline.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2_kernel|wgrad_tc_kernel" -s 3 -c 9 -f -o gpurun_out/r03_ncu_full_tc python tools/ncu_harness.py > gpurun_out/r03_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -s 1400 -c 430 --csv --log-file gpurun_out/r03_step_metrics_ncu.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r03_ncu_list.log 2>&1
python tools/kernel_table.py gpurun_out/r03_step_metrics_ncu.csv > gpurun_out/r03_step_metrics_ncu_summary.md; head -12 gpurun_out/r03_step_metrics_ncu_summary.md; ls -la gpurun_out/r03_ncu_full_tc.ncu-rep
cut -c1-300 gpurun_out/r03_bench_s_b64_n1.json; tail -3 gpurun_out/r03_graph_profile_step.md
