"""One image per call through the host API (the reference's call pattern, Frame::ExtractORB): wall time per call.
Under `ncu --metrics gpu__time_duration.sum` it gives the launch list of a single-image extraction.
  python tools/per_frame_probe.py [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orb_slam_birdview_b200 as pkg          # noqa: E402
from importlib import import_module            # noqa: E402

synth = import_module("orb_slam_birdview_b200.synth")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for (h, w, nf) in ((480, 752, 1000), (376, 1241, 2000), (1080, 1920, 4000)):
    img = synth.synth_frame(h, w, 1000 + h)
    ex = pkg.ORBextractor(nf, 1.2, 8, 20, 7, max_size=(w, h))
    for _ in range(3):
        k, d = ex(img)
    t0 = time.perf_counter()
    for _ in range(reps):
        k, d = ex(img)
    dt = (time.perf_counter() - t0) / reps
    print(f"{w}x{h}/{nf}: {dt * 1e3:.3f} ms per call, {len(k)} keypoints", flush=True)
bimg = synth.synth_frame(400, 400, 3101)
mask = np.full((400, 400), 255, np.uint8)
mask[150:250, 170:230] = 0
ctx = pkg.Context(2000, 1.2, 8, 20, 7, 64, 64)
B = pkg.BirdviewORB(ctx, 2000)
for _ in range(3):
    k, d = B(bimg, mask)
t0 = time.perf_counter()
for _ in range(reps):
    k, d = B(bimg, mask)
print(f"birdview 400x400/2000: {(time.perf_counter() - t0) / reps * 1e3:.3f} ms per call, {len(k)} keypoints", flush=True)
