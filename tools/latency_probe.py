"""Per-call latency of the drop-in path on the GPU box: ORBextractor::operator() through the C++ class (cpp/shim_driver, the
TIMING line) for the two per-frame shapes, under the run-time knobs given on the command line.
  python tools/latency_probe.py [reps] [ENV=1,ENV2=1 ...]     e.g.  python tools/latency_probe.py 300 "" ORBB200_NO_PDL=1"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from importlib import import_module            # noqa: E402

synth = import_module("orb_slam_birdview_b200.synth")
reps = sys.argv[1] if len(sys.argv) > 1 else "300"
variants = sys.argv[2:] or [""]
drv = os.path.join(ROOT, "orb-slam-birdview_b200", "cpp", "shim_driver")
if not os.path.exists(drv):
    subprocess.run(["make", "-s", "-C", os.path.dirname(drv), "shim_driver"], check=True)
with tempfile.TemporaryDirectory() as d:
    for (h, w, nf) in ((480, 752, 1000), (376, 1241, 2000)):
        raw = os.path.join(d, f"in_{w}.raw")
        np.ascontiguousarray(synth.synth_frame(h, w, 1000)).tofile(raw)
        for v in variants:
            env = dict(os.environ, ORBB200_SHIM_TIME=reps)
            for kv in filter(None, v.split(",")):
                k, _, val = kv.partition("=")
                env[k] = val
            r = subprocess.run([drv, raw, str(w), str(h), str(nf), "20", "7", os.path.join(d, "out.bin")], capture_output=True, text=True, env=env)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("TIMING ")]
            print(json.dumps({"shape": f"{w}x{h}", "knobs": v, **(json.loads(line[0][7:]) if line else {"error": r.stderr[-300:]})}), flush=True)

# ---- the whole north-star frame, one frame per call (what bench.py reports as configs.per_frame.north_star_frame_ms) ----
if os.environ.get("ORBB200_PROBE_FRAME", "1") != "0":
    import time
    probe = r'''
import os, sys, time, json
import numpy as np
sys.path.insert(0, %r)
import orb_slam_birdview_b200 as pkg
from importlib import import_module
synth = import_module("orb_slam_birdview_b200.synth")
W, H, BW, BH = 1241, 376, 400, 400
seq = synth.northstar_sequence(24, 5, w=W, h=H, bird=(BW, BH))
ex = pkg.ORBextractor(2000, 1.2, 8, 20, 7, max_size=(W, H))
mp = synth.northstar_map(seq, lambda im: ex(im), 3000, 12)
ctx = pkg.Context(2000, 1.2, 8, 20, 7, W, H, 2)
M = pkg.LocalMap(ctx, mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"], mp["desc"])
step = pkg.FrameStep(ctx, W, H, M, mb=0.537, mbf=386.1448, th=1.0, nnratio=0.8, bird_size=(BW, BH), bird_nfeatures=2000,
                     bird_mask=seq["bird_mask"], bird_window=15, bird_nnratio=0.99)
poses = [pkg.CameraPose.make(**p) for p in seq["poses"]]
n = len(poses)
for i in range(6):
    step(seq["imgs"][2 * i:2 * i + 2], seq["bird_imgs"][i:i + 1], poses[i:i + 1], chain=i > 0)
t0 = time.perf_counter()
for i in range(6, n):
    step(seq["imgs"][2 * i:2 * i + 2], seq["bird_imgs"][i:i + 1], poses[i:i + 1], chain=True)
ms = (time.perf_counter() - t0) / (n - 6) * 1e3
print(json.dumps({"north_star_frame_ms": round(ms, 4)}))
''' % ROOT
    for v in variants:
        env = dict(os.environ)
        for kv in filter(None, v.split(",")):
            k, _, val = kv.partition("=")
            env[k] = val
        r = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True, env=env)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        print(json.dumps({"shape": "north-star frame", "knobs": v, **(json.loads(line[-1]) if line else {"error": r.stderr[-400:]})}), flush=True)
