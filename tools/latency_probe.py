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
