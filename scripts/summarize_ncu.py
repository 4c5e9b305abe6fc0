"""Summarises Nsight Compute captures brought back in gpurun_out/ into small text files for
profiles/ (the .ncu-rep files themselves are too large to commit).

    python scripts/summarize_ncu.py gpurun_out/<tag> profiles/r01/<tag>
reads <tag>_trace.ncu-rep, <tag>_focus.ncu-rep and <tag>_launches.csv."""

import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def raw_page(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    return [dict(zip(header, zip(row, units))) for row in rows[2:]]


def summarise(report, title):
    lines = [f"== {title} ({report})"]
    for launch in raw_page(report):
        name = launch.get("Kernel Name", ("?", ""))[0]
        lines.append(f"kernel: {name}")
        for key in KEYS:
            if key in launch:
                value, unit = launch[key]
                lines.append(f"  {key:84s} {value} {unit}")
    return "\n".join(lines)


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    totals = {}
    for r in rows:
        # ID, PID, process, host, kernel, context, stream, block, grid, device, cc, section, metric, unit, value
        name, value = r[4].split("(")[0], float(r[-1].replace(",", ""))
        count, total = totals.get(name, (0, 0.0))
        totals[name] = (count + 1, total + value)
    grand = sum(t for _, t in totals.values()) or 1.0
    unit = rows[0][-2] if rows else "?"
    lines = [f"== launch list ({path}): gpu__time_duration.sum, cold-cache serialised replay",
             f"{'kernel':40s} {'launches':>8s} {'total ' + unit:>16s} {'share':>8s}"]
    for name, (count, total) in sorted(totals.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{name:40s} {count:8d} {total:16.1f} {total / grand:8.2%}")
    return "\n".join(lines)


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    import os

    parts = [launches(src + "_launches.csv")]
    for suffix, title in (("_trace.ncu-rep", "tracer, --set full"), ("_focus.ncu-rep", "focus stencil, --set full")):
        if os.path.exists(src + suffix):
            parts.append(summarise(src + suffix, title))
    with open(dst + "_ncu_summary.txt", "w") as f:
        f.write("\n\n".join(parts) + "\n")
    print("\n\n".join(parts))
