"""Stall samples per CUDA source line of one kernel in an ncu report (latency-bound kernels: where the warps WAIT, not what they
execute).  python tools/ncu_stalls.py <report.ncu-rep> <kernel name part> [top N]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
mine = [i for i in starts if kern in rows[i][1] and rows[i - 1][1].endswith(".cu")]
if not mine:
    sys.exit("no such kernel")
blk = mine[0]
end = next((i for i in starts if i > blk), len(rows) + 1) - 1
hdr = rows[blk + 1]
sm, src = hdr.index("# Samples"), hdr.index("Source")
ie = hdr.index("Instructions Executed")
data = []
for r in rows[blk + 2:end]:
    try:
        s = int(r[sm] or 0)
    except (ValueError, IndexError):
        continue
    if r[0].isdigit():
        data.append((s, int(r[0]), int(r[ie] or 0), r[src].strip()[:120]))
tot = sum(d[0] for d in data)
print(f"kernel {rows[blk][1].split('(')[0]}: {tot} samples (lines attributed through inlining count at the helper AND the call site)")
for s, ln, n, text in sorted(data, reverse=True)[:top]:
    print(f"{100.0 * s / max(tot, 1):5.1f} %  line {ln:>5}  {n:>9} inst  {text}")
