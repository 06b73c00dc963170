#!/usr/bin/env python
"""Text summary of an ncu --set full report (run here, no GPU needed):  tools/ncu_summary.py report.ncu-rep [kernel regex]
Prints, per captured launch, the metrics DESIGN.md and bench.py's roofline quote."""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {k: i for i, k in enumerate(hdr)}
WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
    "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_global_ld.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum",
    "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__sass_branch_targets_threads_divergent.sum",
]
for r in data:
    name = r[col["Kernel Name"]] if "Kernel Name" in col else "?"
    if pat and not pat.search(name):
        continue
    print("=" * 100)
    print("kernel:", name[:200])
    for k in WANT:
        if k in col:
            print(f"  {k:82s} {r[col[k]]:>18s} {units[col[k]]}")
    print("  -- warp stall reasons (warps per issue-active cycle, > 0.1) --")
    for k, i in col.items():
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0.1:
                print(f"  {k:82s} {v:18.3f}")
