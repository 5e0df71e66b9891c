"""Per-call latency of ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) and BirdviewMatch through the C++ adapters
(cpp/ORBmatcher_b200.cc over Frame / MapPoint objects, cpp/matcher_driver with ORBB200_MATCHER_TIME): a KITTI-sized case, 2000
keypoints and 3000 local map points, 2000 birdview keypoints per frame.  python tools/cpp_matcher_latency.py [reps]"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases          # noqa: E402

reps = sys.argv[1] if len(sys.argv) > 1 else "200"
drv = os.path.join(ROOT, "orb-slam-birdview_b200", "cpp", "matcher_driver")
rng = np.random.default_rng(77)
w, h, nF, nq, th, ratio = 1241, 376, 2000, 3000, 1.0, 0.8
kps, desc, uR, grid = cases.frame_case(nF, w, h, 71, stereo_frac=0.4)
q = cases.projection_queries(kps, desc, uR, w, h, nq, 72)
kp_obs = np.where(rng.random(nF) < 0.1, 3, np.where(rng.random(nF) < 0.1, 0, -1)).astype(np.int32)
bad = ((rng.random(nq) < 0.03) & (q["valid"] == 1)).astype(np.uint8)
q_obs = np.where(q["obs_pos"] == 1, 2, 0).astype(np.int32)
(k1, d1), (k2, d2), bgrid = cases.bird_pair(2000, 400, 73)
hasmp1 = (rng.random(len(k1)) < 0.7).astype(np.uint8)
with tempfile.TemporaryDirectory() as d:
    case, out = os.path.join(d, "case.bin"), os.path.join(d, "out.bin")
    with open(case, "wb") as f:
        np.array([nF, nq, len(k1), len(k2), 15], np.int32).tofile(f)
        np.array([th, ratio, 0.99, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"], bgrid["inv_w"], bgrid["inv_h"]], np.float32).tofile(f)
        for a in (kps, desc, uR, kp_obs, q["valid"], bad, q["u"], q["v"], q["uR"], q["viewcos"], q["level"], q_obs, q["desc"], k1, d1, hasmp1, k2, d2):
            np.ascontiguousarray(a).tofile(f)
    r = subprocess.run([drv, case, out], capture_output=True, text=True, env=dict(os.environ, ORBB200_MATCHER_TIME=reps))
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("TIMING ")]
    print(json.dumps({"knobs": {k: v for k, v in os.environ.items() if k.startswith("ORBB200_")}, **(json.loads(line[0][7:]) if line else {"error": r.stderr[-300:]})}))
