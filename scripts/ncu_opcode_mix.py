"""Executed warp instructions per opcode (and per pipe class) of one kernel, per 32 pixel-samples,
from an `ncu --set full --import-source on` report.

    python scripts/ncu_opcode_mix.py gpurun_out/x.ncu-rep <envs> [H] [spp]
"""
import collections
import csv
import io
import subprocess
import sys

ALU = {"LOP3", "SHF", "IADD3", "ISETP", "FSETP", "PLOP3", "SEL", "FSEL", "FMNMX", "PRMT", "VIADD", "LEA",
       "IABS", "MOV", "BREV", "FLO", "POPC", "IADD", "VIMNMX", "LOP", "IMNMX", "R2P", "P2R", "VOTE", "FCHK"}
FMA = {"FFMA", "FMUL", "FADD", "IMAD", "HADD2", "HFMA2", "FFMA2", "FMUL2", "FADD2"}
XU = {"MUFU", "I2F", "F2F", "F2I", "I2FP"}
CTRL = {"BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "NOP"}
LSU = {"LDS", "STS", "LDG", "STG", "LDL", "STL", "LDC", "LDCU"}


def main():
    report, envs = sys.argv[1], int(sys.argv[2])
    height = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    spp = int(sys.argv[4]) if len(sys.argv) > 4 else 100
    text = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--print-source", "sass"],
                          capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header = next(r for r in rows if r and r[0] == "Address")
    i_src, i_inst, i_thr = header.index("Source"), header.index("Instructions Executed"), header.index(
        "Thread Instructions Executed")
    ops, threads = collections.Counter(), collections.Counter()
    for r in rows:
        if len(r) <= i_thr or not r[0].startswith("0x"):
            continue
        tokens = r[i_src].split()
        op = tokens[1] if tokens[0].startswith("@") else tokens[0]
        op = op.rstrip(";")
        base = op.split(".")[0]
        if base == "IMAD" and (".WIDE" in op or ".HI" in op):
            base = op
        if base in ("I2F", "F2F", "F2I"):
            base = ".".join(op.split(".")[:3])
        ops[base] += int(r[i_inst])
        threads[base] += int(r[i_thr])
    units = envs * height * height * spp / 32.0
    total = sum(ops.values())
    print(f"kernel report {report}: {total / units:.1f} warp instructions per 32 pixel-samples")
    classes = collections.Counter()
    for name, count in ops.most_common():
        key = name.split(".")[0]
        cls = ("alu" if key in ALU else "fma" if key in FMA else "xu" if key in XU else "ctrl" if key in CTRL
               else "lsu" if key in LSU else "fp64" if key in ("DFMA", "DMUL", "DADD") else "other")
        classes[cls] += count
        if count / units >= 0.05:
            print(f"  {name:18s} {count / units:7.2f}  lanes {threads[name] / max(count, 1):5.1f}  [{cls}]")
    print("per class:", {k: round(v / units, 1) for k, v in classes.most_common()})


if __name__ == "__main__":
    main()
