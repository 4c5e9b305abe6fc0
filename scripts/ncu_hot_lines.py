"""Per-source-line share of executed warp instructions and stall samples of one kernel, from
an `ncu --set full --import-source on` report (read offline with `ncu -i ... --page source`).

    python scripts/ncu_hot_lines.py gpurun_out/r01v10_trace.ncu-rep > profiles/r01/v10_mc7_hot_lines.md
"""

import collections
import csv
import io
import os
import subprocess
import sys


def _number(text):
    try:
        return int(text)
    except (TypeError, ValueError):
        return 0


def main():
    report = sys.argv[1]
    text = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                          capture_output=True, text=True, check=True).stdout
    lines = collections.OrderedDict()
    current_file, header, kernel = None, None, None
    for row in csv.reader(io.StringIO(text)):
        if not row:
            continue
        if row[0] == "File Path":
            current_file = os.path.basename(row[1])
        elif row[0] == "Function Name":
            kernel = row[1]
        elif row[0] == "Line No":
            header = row
        elif header and row[0].isdigit():  # a CUDA-C source line (SASS rows have an empty first column)
            record = dict(zip(header, row))
            key = (current_file, int(row[0]))
            entry = lines.setdefault(key, {"source": row[1].strip(), "inst": 0, "samples": 0})
            entry["inst"] += _number(record.get("Instructions Executed"))
            entry["samples"] += _number(record.get("# Samples"))
    total_inst = sum(e["inst"] for e in lines.values()) or 1
    total_samples = sum(e["samples"] for e in lines.values()) or 1
    print(f"# Hot source lines of `{kernel}`\n")
    print(f"From `{os.path.basename(report)}` (ncu --set full, 256 envs x 300x300 x 100 spp): share of executed "
          "warp instructions and of warp-stall samples per CUDA-C line, top 40.\n")
    print("| file:line | instr % | samples % | source |")
    print("|---|---:|---:|---|")
    for (name, number), e in sorted(lines.items(), key=lambda kv: -kv[1]["inst"])[:40]:
        source = e["source"].replace("|", "\\|")[:90]
        print(f"| {name}:{number} | {100 * e['inst'] / total_inst:.1f} | {100 * e['samples'] / total_samples:.1f} "
              f"| `{source}` |")
    by_file = collections.Counter()
    for (name, _), e in lines.items():
        by_file[name] += e["inst"]
    print("\n| file | instr % |\n|---|---:|")
    for name, inst in by_file.most_common():
        print(f"| {name} | {100 * inst / total_inst:.1f} |")


if __name__ == "__main__":
    main()
