"""Wall time of one ORBmatcher call through the C ABI (host arrays in, host arrays out; the frame is device-resident, as the C++
adapters cache it): SearchByProjection(F, 3000 map points) and BirdviewMatch-sized searches.  python tools/matcher_latency.py [reps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orb_slam_birdview_b200 as pkg          # noqa: E402
from importlib import import_module            # noqa: E402

synth = import_module("orb_slam_birdview_b200.synth")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
W, H = 1241, 376
img = synth.synth_frame(H, W, 77)
ex = pkg.ORBextractor(2000, 1.2, 8, 20, 7, max_size=(W, H))
k, d = ex(img)
grid = (0.0, 0.0, 64.0 / W, 48.0 / H)
F = pkg.Frame(ex.ctx, k, d, *grid)
rng = np.random.default_rng(5)
nq = 3000
idx = rng.integers(0, len(k), nq)
args = (np.ones(nq, np.uint8), (k["x"][idx] + rng.uniform(-2, 2, nq)).astype(np.float32), (k["y"][idx] + rng.uniform(-2, 2, nq)).astype(np.float32),
        np.full(nq, -1, np.float32), k["octave"][idx].astype(np.int32), np.full(nq, 0.9, np.float32), synth.perturb_descriptors(d[idx].copy(), 20, 3))
M = pkg.ORBmatcher(ex.ctx, 0.8)
for _ in range(5):
    nm, *_ = M.SearchByProjection(F, *args)
t0 = time.perf_counter()
for _ in range(reps):
    nm, *_ = M.SearchByProjection(F, *args)
ms = (time.perf_counter() - t0) / reps * 1e3
print(json.dumps({"call": "SearchByProjection(F, 3000 map points), frame resident", "ms_per_call": round(ms, 4), "matches": int(nm),
                  "knobs": {k_: v for k_, v in os.environ.items() if k_.startswith("ORBB200_")}}))
