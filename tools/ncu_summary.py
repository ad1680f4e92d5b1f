#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text table: one block per captured launch."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
        ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("sm__inst_executed.avg.per_cycle_active", "IPC (per SM)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("sm__pipe_tensor_subunit_throughput.avg.pct_of_peak_sustained_active", "tensor pipe % (subunit)"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe inst"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active %"),
        ("sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active", "tensor subunit cycles active %"),
        ("sm__inst_executed_pipe_tensor_subunit.sum", "tensor-subunit inst"),
        ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe inst"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("l1tex__data_bank_conflicts_pipe_lsu.sum", "L1/smem bank conflicts"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar / issue"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving / issue"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch / issue"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
        ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping / issue")]
units = rows[1]
for r in rows[2:]:
    print("-" * 100)
    for key, label in want:
        if key in h:
            i = h.index(key)
            print(f"{label:34s} {r[i][:120]} {units[i] if key != 'Kernel Name' else ''}")
