#!/bin/bash
# counters of the shipped kernels at bench size (metrics-only passes; the sass_thread_inst metrics instrument the code: ~100 s for C3)
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed
mkdir -p gpurun_out
ncu --metrics $M --clock-control none -k regex:mega --launch-skip 2 -c 1 --csv --log-file gpurun_out/r2_counters_C3.csv python dev/prof_render.py box_mirror 1920 1080 1024 2 3 > gpurun_out/r2_counters_C3.log 2>&1
ncu --metrics $M --clock-control none -k regex:mega --launch-skip 2 -c 1 --csv --log-file gpurun_out/r2_counters_C2.csv python dev/prof_render.py box 1024 768 256 2 3 > gpurun_out/r2_counters_C2.log 2>&1
ncu --metrics $M --clock-control none -k regex:mega --launch-skip 2 -c 1 --csv --log-file gpurun_out/r2_counters_C4.csv python dev/prof_render.py dof_glass 3840 2160 256 2 3 > gpurun_out/r2_counters_C4.log 2>&1
ncu --metrics $M --clock-control none -k regex:mega --launch-skip 1 -c 1 --csv --log-file gpurun_out/r2_counters_C5.csv python dev/prof_render.py spheres10k 1920 1080 64 2 2 > gpurun_out/r2_counters_C5.log 2>&1
tail -n 1 gpurun_out/r2_counters_C5.log
