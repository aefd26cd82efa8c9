"""One row per captured kernel from an .ncu-rep: the metrics DESIGN.md / profiles/ quote.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.csv"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    cols = [c for c in WANT if c in hdr] + stalls
    out = csv.writer(sys.stdout)
    out.writerow([c.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "") for c in cols])
    out.writerow([units[hdr.index(c)] for c in cols])
    for r in rows[2:]:
        out.writerow([r[hdr.index(c)] for c in cols])


if __name__ == "__main__":
    main()
