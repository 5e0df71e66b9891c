"""Turn ncu outputs into the markdown tables kept under profiles/.
  python tools/ncu_summary.py launches <launch_list.csv>          -> per-kernel share table
  python tools/ncu_summary.py raw <report.ncu-rep> [metric ...]   -> one column per captured launch
  python tools/ncu_summary.py source <report.ncu-rep> <kernel name part> [min share %]
                                                                  -> executed warp-instructions per opcode and per source line
                                                                     (first captured launch of that kernel; needs -lineinfo + --import-source on)"""
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


def source(path, kernel, min_share=0.5, launch=0):
    """Opcode table from the SASS view (every instruction once); line table from the CUDA view (an instruction inlined from a
    helper is attributed to the helper's line AND to the call site's line there, so line shares are normalised by the SASS total
    and can add up to more than 100 %)."""
    from collections import Counter
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    mine = [i for i in starts if kernel in rows[i][1]]
    if not mine:
        sys.exit(f"no kernel matching {kernel!r} in {path}")
    blk = mine[min(launch, len(mine) - 1)]
    end = next((i for i in starts if i > blk), len(rows))
    hdr = rows[blk + 1]
    ie, sm, te, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed"), hdr.index("Source")
    ops, osmp = Counter(), Counter()
    tot = tsm = tthr = 0
    for r in rows[blk + 2:end]:
        try:
            v = int(r[ie])
        except (ValueError, IndexError):
            continue
        w = r[src].split()
        if not w:
            continue
        op = (w[1] if w[0].startswith("@") else w[0]).split(".")[0]
        ops[op] += v
        osmp[op] += int(r[sm] or 0)
        tot += v
        tsm += int(r[sm] or 0)
        tthr += int(r[te] or 0)
    print(f"kernel `{rows[blk][1].split('(')[0]}`: {tot} warp-instructions executed, {tthr / (32.0 * tot):.3f} of the lanes active on average\n")
    print("| opcode | share of executed warp-instructions | share of stall samples |\n|---|---|---|")
    for op, v in ops.most_common():
        if 100.0 * v / tot >= min_share:
            print(f"| {op} | {100.0 * v / tot:.1f} % | {100.0 * osmp[op] / max(tsm, 1):.1f} % |")
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
    mine = [i for i in starts if kernel in rows[i][1] and rows[i - 1][1].endswith(".cu")]
    if not mine:
        return
    blk = mine[0]
    end = next((i for i in starts if i > blk), len(rows) + 1) - 1
    hdr = rows[blk + 1]
    ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
    ws, wi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
    lines = []
    for r in rows[blk + 2:end]:
        if len(r) <= ie or r[0] == "":
            continue
        try:
            lines.append((int(r[0]), int(r[ie]), int(r[sm] or 0), int(r[ws] or 0), int(r[wi] or 0), r[1].strip()[:96]))
        except ValueError:
            continue
    nl = max(len(mine), 1)      # the CUDA view lists the function once per captured launch? no: once per file; counts are of one launch
    print("\n| line | warp-instructions (share of the kernel's total) | stall samples | shared wavefronts (ideal) | source |\n|---|---|---|---|---|")
    ls = sum(x[2] for x in lines) or 1
    tw = sum(x[3] for x in lines) or 1
    for ln, v, smp, w, wid, text in sorted(lines):
        if 100.0 * v / tot >= min_share or w > 0.03 * tw:
            print(f"| {ln} | {100.0 * v / tot:.1f} % | {100.0 * smp / ls:.1f} % | {w / 1e6:.1f} M ({wid / 1e6:.1f} M) | `{text}` |")


if __name__ == "__main__":
    if sys.argv[1] == "source":
        source(sys.argv[2], sys.argv[3], float(sys.argv[4]) if len(sys.argv) > 4 else 0.5)
        sys.exit(0)
    {"launches": lambda: launches(sys.argv[2]), "raw": lambda: raw(sys.argv[2], sys.argv[3:])}[sys.argv[1]]()
