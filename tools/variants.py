"""Build experiment variants of liborbb200.so into variants/<name>.so (git-ignored, shipped to the GPU box).
usage: python tools/variants.py name1:"-DFOO=1 -DBAR=2" name2:""   then on the box: tools/run_variants.sh"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "orb-slam-birdview_b200", "csrc")
SOURCES = ["api.cu", "extract.cu", "match.cu", "stereo.cu", "bow.cu", "bird.cu"]
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    out = os.path.join(ROOT, "variants", name + ".so")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared", "-o", out] + flags.split() + SOURCES
    procs.append((name, subprocess.Popen(cmd, cwd=CSRC)))
for name, p in procs:
    if p.wait() != 0:
        sys.exit(f"variant {name} failed to build")
    print("built", name)
