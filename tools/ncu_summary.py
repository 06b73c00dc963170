import csv, sys, subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__waves_per_multiprocessor','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__cycles_active.avg','sm__cycles_elapsed.avg','sm__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.sum',
'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__sass_average_branch_targets_threads_uniform.pct','smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr: print(f"  {w:82s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
if '-a' in sys.argv:
    for h in hdr:
        if 'pipe' in h and ('pct' in h): print(h, rows[2][hdr.index(h)])
