"""Condenses an `ncu --set full` report into the per-launch text summary kept under profiles/.
usage: python scripts/ncu_summary.py gpurun_out/r1d_prof.ncu-rep > profiles/r1d_ncu_full.txt
(runs `ncu -i <rep> --page raw --csv` here; no GPU needed)"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[head], rows[head + 1]
    col = {n: i for i, n in enumerate(names)}
    print(f"ncu --set full --clock-control none, {rep}; one block per profiled launch")
    for k, r in enumerate(rows[head + 2:]):
        if len(r) < len(names):
            continue
        print(f"== launch {k}: {r[col['Kernel Name']][:72]}")
        for m in KEEP:
            if m in col:
                print(f"{m} = {r[col[m]]} {units[col[m]]}")


if __name__ == "__main__":
    main()
