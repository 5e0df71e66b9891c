"""Per-kernel SASS evidence for profiles/sass_summary.txt: which of the instructions the design relies on each kernel of
liborbb200.so contains (cuobjdump -sass; static counts, not executed counts).
  python tools/sass_summary.py [path/to/liborbb200.so] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "orb-slam-birdview_b200", "liborbb200.so")
WATCH = ["UTMALDG", "UTMASTG", "SYNCS", "VIMNMX3", "VIMNMX", "IDP", "POPC", "PRMT", "SHF", "DMUL", "DADD", "DFMA", "F2F", "LDGSTS", "REDUX", "VOTE",
         "ATOMS", "ATOMG", "RED", "BAR", "ACQBULK", "PREEXIT", "HMMA", "UTCHMMA"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        cur["_total"] += 1
        cur[m.group(1)] += 1

demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
print(f"{os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, cubin architectures {arch}")
print("static SASS instruction counts per kernel (cuobjdump -sass); only the watched mnemonics that occur are listed")
print("UTMALDG = cp.async.bulk.tensor (TMA load), SYNCS = mbarrier, ACQBULK / PREEXIT = griddepcontrol.wait / .launch_dependents (programmatic dependent launch), VIMNMX3 = three-input packed min/max, IDP = dp2a/dp4a, no HMMA/UTC*MMA: nothing on this path is a matrix product\n")
for mangled, name in zip(kernels, demangle):
    c = kernels[mangled]
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "").replace("orbb200::", "")
    hits = ", ".join(f"{k} {c[k]}" for k in WATCH if c[k])
    print(f"{short:<44} {c['_total']:>5} instr   {hits}")
