"""Turn ncu outputs into the markdown tables kept under profiles/.
  python tools/ncu_summary.py launches <launch_list.csv>          -> per-kernel share table
  python tools/ncu_summary.py raw <report.ncu-rep> [metric ...]   -> one column per captured launch"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

DEFAULT = ["launch__grid_size", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r]
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void orbb200::", "").replace("orbb200::", "")
        t = float(r[mv].replace(",", "")) / 1000.0      # ns -> us
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"{sum(v[0] for v in agg.values())} launches, {tot:.0f} us total\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % | {v[1] / v[0]:.1f} |")


def raw(path, metrics):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    names = [r[kn].split("(")[0].replace("void orbb200::", "").replace("orbb200::", "") for r in data]
    print("| metric | unit | " + " | ".join(names) + " |\n|---|---|" + "---|" * len(names))
    for m in metrics or DEFAULT:
        if m in hdr:
            i = hdr.index(m)
            short = m.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", "")
            print(f"| {short} | {units[i]} | " + " | ".join(r[i][:10] for r in data) + " |")


if __name__ == "__main__":
    {"launches": lambda: launches(sys.argv[2]), "raw": lambda: raw(sys.argv[2], sys.argv[3:])}[sys.argv[1]]()
