#!/bin/bash
# ncu counters of the two GP kernels of the TF32 train step (tools/prof_step.py first runs clean)
M="gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__waves_per_multiprocessor,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"
python tools/prof_step.py cfg2 3 tf32 > gpurun_out/ps.log 2>&1 || exit 1
ncu --metrics $M --clock-control none --kernel-name regex:gp_.*_warp_kernel -c 6 --csv --log-file gpurun_out/r2_gp_counters.csv python tools/prof_step.py cfg2 3 tf32 > gpurun_out/ncu_gp.log 2>&1
ls -la gpurun_out/r2_gp_counters.csv
