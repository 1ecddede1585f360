#!/bin/bash
# key counters of the tier kernels for a variant (1 M pairs): IPC, lanes, instructions, smem conflicts, stalls
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
for name in "$@"; do
  if [ "$name" = default ]; then unset CB2_LIB_PATH; else export CB2_LIB_PATH=$PWD/collision_rs_b200/variants/libcb2_$name.so; fi
  CB2_PROFILE_PAIRS=1000000 timeout 600 ncu --metrics $M --clock-control none -k regex:k_epa --launch-skip 3 --launch-count 2 --csv --log-file gpurun_out/n_$name.csv python tools/profile_pass.py cfg3 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/n_$name.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
cur=None
for r in rows[1:]:
    if r[ii]!=cur: cur=r[ii]; print('== $name', r[ki][:50])
    print('   %-95s %s'%(r[mi].replace('smsp__average_warps_issue_stalled_','stall_'), r[vi]))
PY
done
