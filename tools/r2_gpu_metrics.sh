#!/bin/bash
# per-kernel counters of every config at full size (no replays of --set full: a fixed metric list)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed_pipe_xu.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
timeout 600 python tools/profile_pass.py > gpurun_out/m_plain.log 2>&1 || { echo "plain pass failed"; tail -5 gpurun_out/m_plain.log; exit 1; }
timeout 1500 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_metrics.csv python tools/profile_pass.py > gpurun_out/m_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2_metrics.csv; tail -2 gpurun_out/m_ncu.log
