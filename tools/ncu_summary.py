#!/usr/bin/env python
"""Per-launch summary of `ncu --set full` reports: python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...] > profiles/x.txt
(the .ncu-rep files stay in gpurun_out/, scratch; what is committed are these extracted numbers)."""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (hmma) active %"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU wavefronts %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio throttle / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg throttle / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected / issue"),
]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print("== %s: no data" % rep)
        continue
    h, u = rows[0], rows[1]
    print("== %s" % rep.split("/")[-1])
    for r in rows[2:]:
        print("  -- %s  (launch id %s)" % (r[h.index("Kernel Name")][:110], r[h.index("ID")]))
        for key, label in WANT:
            if key in h and r[h.index(key)] not in ("", "0", "0.000000"):
                print("     %-34s %s %s" % (label, r[h.index(key)], u[h.index(key)]))
    print()
