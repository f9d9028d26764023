"""Print selected metrics from an `ncu --page raw --csv` dump: python tools/ncu_pick.py raw.csv [substr ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
default = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__average_warp", "smsp__average_warps_issue_stalled", "dram__cycles_active", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum"]
pats = sys.argv[2:] or default
for i, h in enumerate(hdr):
    if any(p in h for p in pats):
        print("%-90s %-14s %s" % (h, units[i], [r[i] for r in data]))
