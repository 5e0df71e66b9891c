#!/usr/bin/env python3
"""Golden vectors of the birdview front-end (reference src/Frame.cc:328-342) from the real OpenCV (cv2 4.13):

    orb = cv2.ORB_create(2000); kps = orb.detect(img, mask); cv2.cornerSubPix(img, pts, (5,5), (-1,-1), (EPS+ITER,40,1e-3));
    kps, desc = orb.compute(img, kps)

Run in the build container only:  python tests/golden/make_golden_bird.py
cv2.setUseOptimized(False): OpenCV's own C++ code paths.  With IPP enabled cv2.getRectSubPix (inside cornerSubPix)
interpolates with a different float operation order (last-ulp differences in the refined corners); ORB detect/compute do
not depend on the switch.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import cv2  # noqa: E402

import cases  # noqa: E402
from helpers import KP_DTYPE  # noqa: E402

cv2.setNumThreads(1)
cv2.setUseOptimized(False)


def kp_array(kps):
    a = np.empty(len(kps), KP_DTYPE)
    for i, k in enumerate(kps):
        a[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
    return a


def front_end(img, mask, nfeatures=2000):
    orb = cv2.ORB_create(nfeatures)
    det = orb.detect(img, mask)
    pts = np.array([k.pt for k in det], np.float32).reshape(-1, 1, 2)
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 40, 0.001)
    sub = cv2.cornerSubPix(img, pts.copy(), (5, 5), (-1, -1), crit).reshape(-1, 2) if len(det) else pts.reshape(-1, 2)
    moved = list(det)
    for k, p in zip(moved, sub):
        k.pt = (float(p[0]), float(p[1]))
    d0 = kp_array(det)          # (pt already replaced: cv2.KeyPoint objects are shared) -> rebuild the detect record
    det_arr = d0.copy()
    det_arr["x"], det_arr["y"] = pts[:, 0, 0], pts[:, 0, 1]
    kept, desc = orb.compute(img, moved)
    return det_arr, sub, kp_array(kept), (desc if desc is not None else np.zeros((0, 32), np.uint8))


def main():
    for name, (size, seed, with_mask) in {"400": (400, 3101, True), "384_nomask": (384, 3102, False), "500x360": ((500, 360), 3103, True)}.items():
        img, mask = cases.birdview_case(size, seed)
        det, sub, kept, desc = front_end(img, mask if with_mask else None)
        # getRectSubPix samples incl. windows that leave the image
        rng = np.random.default_rng(seed)
        h, w = img.shape
        centers = np.stack([rng.uniform(-8, w + 8, 300), rng.uniform(-8, h + 8, 300)], 1).astype(np.float32)
        centers[:40] = np.round(centers[:40])
        patches = np.stack([cv2.getRectSubPix(img, (13, 13), (float(c[0]), float(c[1])), patchType=cv2.CV_32F) for c in centers])
        np.savez_compressed(os.path.join(HERE, f"bird_orb_{name}.npz"), size=np.array(img.shape[::-1], np.int32), seed=np.int32(seed),
                            with_mask=np.int32(with_mask), detect=det, subpix=sub, kps=kept, desc=desc, centers=centers, patches=patches)
        print(name, len(det), len(kept))
    # INTER_LINEAR_EXACT resize
    out = {}
    rng = np.random.default_rng(9)
    for i, (w, h, dw, dh) in enumerate([(200, 200, 167, 167), (167, 167, 139, 139), (97, 61, 81, 51), (150, 108, 125, 90)]):
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        out[f"src{i}"] = src
        out[f"dst{i}"] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT)
    src = rng.integers(0, 256, (120, 160), dtype=np.uint8)
    g = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    out["sep_src"] = src
    out["sep_dst"] = cv2.sepFilter2D(src, cv2.CV_8U, g, g, borderType=cv2.BORDER_REFLECT_101)
    np.savez_compressed(os.path.join(HERE, "bird_primitives.npz"), **out)


if __name__ == "__main__":
    main()
