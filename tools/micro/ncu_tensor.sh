#!/bin/bash
# ncu tensor-pipe / DRAM counters of the tcgen05 GEMM launches of the full-batch step and the eval pass (each program first runs clean)
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active"
for p in tf32 bf16; do
  python tools/prof_fullbatch.py $p > gpurun_out/fb_$p.log 2>&1 || exit 1
  ncu --metrics $M --clock-control none --kernel-name regex:tc_gemm -c 10 --csv --log-file gpurun_out/r2_tensor_fullbatch_$p.csv python tools/prof_fullbatch.py $p > gpurun_out/ncu_fb_$p.log 2>&1
done
python tools/prof_eval.py tf32 collapsed 2 > gpurun_out/ev.log 2>&1 || exit 1
ncu --metrics $M --clock-control none --kernel-name regex:tc_gemm -c 6 --csv --log-file gpurun_out/r2_tensor_eval_tf32.csv python tools/prof_eval.py tf32 collapsed 2 > gpurun_out/ncu_ev.log 2>&1
python tools/prof_eval.py bf16 materialised 2 > gpurun_out/ev2.log 2>&1 || exit 1
ncu --metrics $M --clock-control none --kernel-name regex:tc_gemm -c 6 --csv --log-file gpurun_out/r2_tensor_eval_bf16_mat.csv python tools/prof_eval.py bf16 materialised 2 > gpurun_out/ncu_ev2.log 2>&1
ls -la gpurun_out/r2_tensor_*.csv
