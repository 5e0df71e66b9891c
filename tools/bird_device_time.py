#!/usr/bin/env python3
"""Device time of the birdview front-end per batch (CUDA events on the context's stream around the host-API call minus
nothing: the call's H2D/D2H are inside; plus a device-only figure from the per-kernel ncu list when run under ncu).
Prints wall ms per call for a few batch sizes."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
import orb_slam_birdview_b200 as pkg
ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
B = pkg.BirdviewORB(ctx, 2000)
_, mask = cases.birdview_case(400, 3101)
out = {}
for batch in (1, 8, 64):
    imgs = [cases.birdview_case(400, 4000 + i)[0] for i in range(batch)]
    masks = [mask] * batch
    B.extract_batch(imgs, masks)
    t = time.perf_counter()
    for _ in range(3):
        B.extract_batch(imgs, masks)
    out[batch] = (time.perf_counter() - t) / 3 * 1e3
print(json.dumps(out))
