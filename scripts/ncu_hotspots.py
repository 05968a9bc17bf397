"""Per-opcode breakdown of executed instructions and stall samples from an .ncu-rep captured with --import-source on
(SASS view of the source page).   python scripts/ncu_hotspots.py gpurun_out/prof_fe3.ncu-rep"""
import collections
import csv
import subprocess
import sys


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head = rows[1]
    i_src, i_smp, i_exe = head.index("Source"), head.index("# Samples"), head.index("Instructions Executed")
    ins, smp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= i_exe:
            continue
        op = r[i_src].split()
        if not op:
            continue
        name = op[1] if op[0].startswith("@") else op[0]
        name = name.split(".")[0] + (".WIDE" if ".WIDE" in name else "")
        ins[name] += int(r[i_exe] or 0)
        smp[name] += int(r[i_smp] or 0)
    ti, ts = sum(ins.values()), sum(smp.values())
    print("kernel,", rows[0][1])
    print("opcode,instructions_executed_pct,stall_samples_pct")
    for k, v in ins.most_common(14):
        print(f"{k},{100.0 * v / ti:.2f},{100.0 * smp[k] / max(ts, 1):.2f}")


if __name__ == "__main__":
    main()
