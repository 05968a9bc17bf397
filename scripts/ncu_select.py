"""Reduce an .ncu-rep (ncu --set full) to the metrics quoted in DESIGN.md / profiles/README.md.
    python scripts/ncu_select.py gpurun_out/prof_dense.ncu-rep > profiles/r01_fe_dense_ncu_raw_selected.csv"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|launch__(registers_per_thread|grid_size|block_size|waves_per_multiprocessor)"
    r"|sm__cycles_elapsed\.avg\.per_second|sm__inst_executed_pipe_(alu|fma|xu|uniform)\.avg\.pct_of_peak_sustained_active"
    r"|sm__pipe_(fmaheavy|fma|alu)_cycles_active\.avg\.pct_of_peak_sustained_elapsed|sm__issue_active\.avg\.pct_of_peak_sustained_elapsed"
    r"|sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active\.ratio"
    r"|smsp__inst_executed\.sum|smsp__thread_inst_executed\.sum|smsp__thread_inst_executed_per_inst_executed\.ratio"
    r"|smsp__sass_average_branch_targets_threads_uniform\.pct|sm__throughput\.avg\.pct_of_peak_sustained_elapsed)$")


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, unit = rows[0], rows[1]
    w = csv.writer(sys.stdout)
    for r in rows[2:]:
        name = r[head.index("Kernel Name")] if "Kernel Name" in head else ""
        w.writerow(["kernel", "", name])
        w.writerow(["metric", "unit", "value"])
        for h, u, v in zip(head, unit, r):
            if KEEP.match(h) and not (h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) < 0.05):
                w.writerow([h, u, v])


if __name__ == "__main__":
    main()
