#!/usr/bin/env python3
"""Birdview front-end (cv::ORB detect + cornerSubPix + compute, reference src/Frame.cc:328-342) timing: the device
pipeline through the C ABI with host buffers vs the CPU oracle and, where importable, cv2 itself.  Side measurement
(C3's birdview half); prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=400)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import cases
    import orb_slam_birdview_b200 as pkg
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    B = pkg.BirdviewORB(ctx, 2000)
    img, mask = cases.birdview_case(a.size, 3101)
    for _ in range(3):
        B(img, mask)
    t = time.perf_counter()
    for _ in range(a.reps):
        k, d = B(img, mask)
    single_ms = (time.perf_counter() - t) / a.reps * 1e3
    imgs = [cases.birdview_case(a.size, 4000 + i)[0] for i in range(a.batch)]
    masks = [mask] * a.batch
    B.extract_batch(imgs, masks)
    t = time.perf_counter()
    for _ in range(a.reps):
        B.extract_batch(imgs, masks)
    batch_ms = (time.perf_counter() - t) / a.reps / a.batch * 1e3
    out = {"workload": f"birdview {a.size}x{a.size}, cv::ORB(2000) detect(mask) + cornerSubPix(5x5,40,1e-3) + compute", "keypoints": int(len(k)),
           "gpu_single_image_ms": single_ms, "gpu_batch_ms_per_image": batch_ms, "batch": a.batch, "host_buffers": True}
    if not a.no_cpu:
        import oracle
        t = time.perf_counter()
        for _ in range(3):
            oracle.bird_extract(img, mask, 2000)
        out["cpu_oracle_ms"] = (time.perf_counter() - t) / 3 * 1e3
        try:
            import cv2
            cv2.setNumThreads(1)
            orb = cv2.ORB_create(2000)
            crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 40, 0.001)
            t = time.perf_counter()
            for _ in range(3):
                kk = orb.detect(img, mask)
                p = np.array([q.pt for q in kk], np.float32).reshape(-1, 1, 2)
                p = cv2.cornerSubPix(img, p, (5, 5), (-1, -1), crit)
                for q, r in zip(kk, p.reshape(-1, 2)):
                    q.pt = (float(r[0]), float(r[1]))
                orb.compute(img, kk)
            out["cv2_1thread_ms"] = (time.perf_counter() - t) / 3 * 1e3
        except ImportError:
            pass
    print(json.dumps(out))


if __name__ == "__main__":
    main()
