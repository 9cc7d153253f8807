"""Condense `ncu --page raw --csv` output to the metrics the roofline discussion uses.
python tools/ncu_summary.py <report.ncu-rep> [kernel regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    units = rows[1]
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if pat and not pat.search(name):
            continue
        print("== %s  grid %s block %s" % (name[:60], r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]))
        for k in KEYS:
            if k in col:
                print("   %-90s %s %s" % (k, r[col[k]], units[col[k]]))


if __name__ == "__main__":
    main()
