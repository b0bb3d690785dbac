#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
M="sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active,sm__cycles_active.avg,sm__cycles_elapsed.avg,gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"
for mode in c4 c4_plain; do
  echo "== range replay: $mode"
  timeout 600 ncu --replay-mode range --clock-control none --metrics $M --csv --log-file $OUT/r3x_range_$mode.csv python scripts/profile_range.py $mode 262144 512 > $OUT/r3x_range_$mode.log 2>&1
  tail -3 $OUT/r3x_range_$mode.log
  grep -v "^==" $OUT/r3x_range_$mode.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r) > 3 and r[-3] != 'Metric Name': print('%-90s %-12s %s' % (r[-3], r[-2], r[-1]))
"
done 2>&1 | tee $OUT/r3x_range.log
