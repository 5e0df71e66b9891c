#!/usr/bin/env python3
"""Generate the committed golden vectors (run in the build container only: needs cv2 4.13).

  python tests/golden/make_golden.py

* primitives.npz  -- cv2 outputs for resize / GaussianBlur / per-cell FAST / fastAtan2 on seeded inputs
* extract_*.npz   -- input image + (keypoints, descriptors) from tests/ref_py/orb_cv2_ref.py, the
                     independent Python composition of cv2 primitives following src/ORBextractor.cc
* matcher.npz     -- outputs of the literal Python transcriptions in tests/ref_py/matcher_py_ref.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "ref_py"))
import cv2  # noqa: E402

import cases  # noqa: E402
import matcher_py_ref as mref  # noqa: E402
from helpers import synth  # noqa: E402
from orb_cv2_ref import OrbCv2Ref  # noqa: E402

cv2.setNumThreads(1)


def primitives():
    out = {}
    rng = np.random.default_rng(42)
    for tag, (h, w) in {"a": (97, 131), "b": (60, 60), "c": (134, 210)}.items():
        img = synth.synth_frame(h, w, 500 + h) if tag != "b" else rng.integers(0, 256, (h, w), dtype=np.uint8)
        dw, dh = int(round(w / 1.2)), int(round(h / 1.2))
        out[f"img_{tag}"] = img
        out[f"resize_{tag}"] = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        out[f"gauss_{tag}"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        for th in (7, 20):
            kps = cv2.FastFeatureDetector_create(th, True).detect(img)
            out[f"fast{th}_{tag}"] = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
        cell = np.ascontiguousarray(img[10:46, 20:57])
        kps = cv2.FastFeatureDetector_create(7, True).detect(cell)
        out[f"fastcell_{tag}"] = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
    yx = rng.integers(-624240, 624241, (4000, 2)).astype(np.float32)
    yx[:8] = [[0, 0], [0, 1], [1, 0], [-1, 0], [0, -1], [1, 1], [-1, -1], [5, -5]]
    out["atan_yx"] = yx
    out["atan_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **out)


def extractor():
    for name, (h, w, nf, ini, mn, seed) in {
        "c1_752x480": (480, 752, 1000, 20, 7, 1000),
        "bird_400x400": (400, 400, 2000, 15, 5, 3001),
        "small_320x240": (240, 320, 500, 20, 7, 77),
    }.items():
        img = synth.synth_frame(h, w, seed)
        k, d = OrbCv2Ref(nf, 1.2, 8, ini, mn)(img)
        np.savez_compressed(os.path.join(HERE, f"extract_{name}.npz"), img=img, kps=k, desc=d,
                            params=np.array([nf, ini, mn, seed], np.int32))
        print(name, len(k))


def matcher():
    out = {}
    w, h = 620, 188
    kps, desc, uR, grid = cases.frame_case(500, w, h, 11, stereo_frac=0.4)
    F = mref.PyFrame(kps, desc, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"], uR)
    q = cases.projection_queries(kps, desc, uR, w, h, 700, 12)
    blocked = (np.random.default_rng(13).random(len(kps)) < 0.1).astype(np.uint8)
    for th in (1.0, 3.0):
        n, qk = mref.search_by_projection(F, cases.SCALE_FACTORS, q["valid"], q["u"], q["v"], q["uR"], q["level"], q["viewcos"],
                                          q["desc"], q["obs_pos"], blocked, th, 0.8)
        out[f"sbp_th{int(th)}"] = np.array([n] + qk, np.int32)
    for mode in (0, 1, 2):
        n, qk = mref.search_by_projection_frame(F, cases.SCALE_FACTORS, q["valid"], q["u"], q["v"], q["invz"], q["level"],
                                                q["angle"], q["desc"], q["obs_pos"], blocked, 7.0, 40.0, mode, True)
        out[f"sbpf_mode{mode}"] = np.array([n] + qk, np.int32)
    (k1, d1), (k2, d2), g = cases.bird_pair(400, 200, 21)
    F2 = mref.PyFrame(k2, d2, g["min_x"], g["min_y"], g["inv_w"], g["inv_h"])
    n, m12, _ = mref.birdview_match(k1, d1, F2, None, 10, 0.99, True)
    out["bird_a"] = np.array([n] + m12, np.int32)
    prev = np.stack([k1["x"], k1["y"]], 1)
    n, m12, prev2 = mref.birdview_match(k1, d1, F2, prev, 15, 0.99, True)
    out["bird_b"] = np.array([n] + m12, np.int32)
    out["bird_b_prev"] = prev2
    has = (np.random.default_rng(22).random(len(k1)) < 0.6).astype(np.uint8)
    n, mk = mref.search_by_match_bird_kf(k1, has, d1, F2, 15.0, 0.99, True)
    out["bird_kf"] = np.array([n] + mk, np.int32)
    qx = k1["x"] + 3
    qy = k1["y"] - 2
    n, qk = mref.search_by_projection_bird(F2, has, qx, qy, d1, None, None, 4.0, 0.99)
    out["bird_proj"] = np.array([n] + qk, np.int32)
    t = cases.triangulation_case(300, 300, 620, 188, 31)

    def todict(fv):
        return {int(n): fv[2][fv[1][i]:fv[1][i + 1]].tolist() for i, n in enumerate(fv[0])}

    for only_stereo in (0, 1):
        n, pairs = mref.search_for_triangulation(t["k1"], t["d1"], t["uR1"], t["has1"], t["k2"], t["d2"], t["uR2"], t["has2"],
                                                 todict(t["fv1"]), todict(t["fv2"]), t["F12"], t["ex"], t["ey"], t["sf2"], t["sigma2"],
                                                 bool(only_stereo), True)
        out[f"tri_{only_stereo}"] = np.array([[n, len(pairs)]] + [list(p) for p in pairs], np.int32)
    np.savez_compressed(os.path.join(HERE, "matcher.npz"), **out)
    print({k: (v[0] if v.ndim == 1 else v[0].tolist()) for k, v in out.items() if k != "bird_b_prev"})


if __name__ == "__main__":
    primitives()
    extractor()
    matcher()
