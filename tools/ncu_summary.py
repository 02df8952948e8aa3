"""Print the roofline-relevant counters of every kernel in an .ncu-rep (ncu --page raw --csv)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct",
        "sm__inst_executed_pipe_tensor", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum ",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "l1tex__throughput.avg.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum ",
        "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum ", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled", "sm__ctas_launched",
        "lts__t_sector_hit_rate.pct", "sm__sass_inst_executed_op_shared", "l1tex__data_bank_conflicts_pipe_lsu.sum "]
for r in rows[2:]:
    print("=" * 100)
    for i, c in enumerate(h):
        cc = c + " "
        if any(k in cc for k in KEYS) and r[i] not in ("", "n/a"):
            print(f"{c:86s} {units[i]:10s} {r[i][:60]}")
