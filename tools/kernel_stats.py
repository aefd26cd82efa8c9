"""profiles/kernel_stats.json from ncu reports: per kernel the DRAM bytes of one launch, the issue-slot
utilisation and the warp instructions per symbol step -- what bench.py puts into `roofline` for the
issue-bound coder kernels (it reads the committed file; it never runs ncu).
    python tools/kernel_stats.py WORKLOAD SYMBOLS report.ncu-rep [more.ncu-rep ...]
also writes profiles/<report>_summary.csv (one row per kernel launch) next to it."""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEEP = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main():
    workload, symbols = sys.argv[1], int(sys.argv[2])
    out_path = ROOT / "profiles" / "kernel_stats.json"
    stats = json.loads(out_path.read_text()) if out_path.exists() else {}
    for rep in sys.argv[3:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        summary = [["kernel"] + KEEP, ["", *[units[col[k]] if k in col else "" for k in KEEP]]]
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            short = name.split("(")[0].split("::")[-1].split("<")[0].replace("void ", "").strip()

            def val(k):
                try:
                    return float(r[col[k]])
                except (KeyError, ValueError):
                    return None

            def bytes_of(k):
                v, u = val(k), units[col[k]] if k in col else ""
                if v is None:
                    return None
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

            summary.append([name] + [r[col[k]] if k in col else "" for k in KEEP])
            inst = val("smsp__inst_executed.sum")
            rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
            stats[f"{workload}:{short}"] = {
                "dram_bytes": None if rd is None else rd + wr,
                "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "inst_per_symbol": None if inst is None else inst * 32.0 / symbols,
                "alu_pipe_pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "fma_pipe_pct": val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
                "ncu_duration_ms": (val("gpu__time_duration.sum") or 0) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(
                    units[col["gpu__time_duration.sum"]], 1),
                "source": f"profiles/{Path(rep).stem}_summary.csv (ncu --set full --clock-control none, one launch; "
                          "inst_per_symbol = warp instructions x 32 / symbols: one warp step codes one symbol of 32 blocks)"}
        with open(ROOT / "profiles" / f"{Path(rep).stem}_summary.csv", "w", newline="") as f:
            csv.writer(f).writerows(summary)
    out_path.write_text(json.dumps(stats, indent=1, sort_keys=True) + "\n")
    print(f"{len(stats)} kernels in {out_path}")


if __name__ == "__main__":
    main()
