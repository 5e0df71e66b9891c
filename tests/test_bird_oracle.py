"""CPU tests pinning the birdview front-end oracle (oracle/bird_oracle.cpp: cv::ORB detect + cornerSubPix + compute as
used by the reference, src/Frame.cc:328-342) against vectors produced by the real OpenCV (tests/golden/bird_*.npz,
made by tests/golden/make_golden_bird.py) and, where cv2 is importable, against live cv2 calls on other inputs."""
import os

import numpy as np
import pytest

import cases
from helpers import GOLDEN, oracle


def _same_kps(a, b):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in a.dtype.names)


@pytest.mark.parametrize("name", ["400", "384_nomask", "500x360"])
def test_bird_front_end_golden(name):
    g = np.load(os.path.join(GOLDEN, f"bird_orb_{name}.npz"))
    w, h = (int(v) for v in g["size"])
    img, mask = cases.birdview_case((w, h), int(g["seed"]))
    mask = mask if int(g["with_mask"]) else None
    det = oracle.bird_detect(img, mask, 2000)
    assert _same_kps(det, g["detect"])                       # positions, Harris responses, angles, sizes AND order
    if mask is not None:
        assert not np.any(mask[np.rint(det["y"] / 1).astype(int)[det["octave"] == 0], np.rint(det["x"]).astype(int)[det["octave"] == 0]] == 0)
    sub = oracle.corner_subpix(img, np.stack([det["x"], det["y"]], 1))
    assert np.array_equal(sub.view(np.uint32), g["subpix"].view(np.uint32))
    moved = det.copy()
    moved["x"], moved["y"] = sub[:, 0], sub[:, 1]
    kps, desc = oracle.bird_compute(img, moved)
    assert _same_kps(kps, g["kps"]) and np.array_equal(desc, g["desc"])
    k2, d2 = oracle.bird_extract(img, mask, 2000)
    assert _same_kps(k2, g["kps"]) and np.array_equal(d2, g["desc"])
    for c, p in zip(g["centers"], g["patches"]):
        assert np.array_equal(oracle.get_rect_sub_pix(img, float(c[0]), float(c[1]), 13, 13), p)


def test_bird_primitives_golden():
    g = np.load(os.path.join(GOLDEN, "bird_primitives.npz"))
    for i in range(4):
        dst = g[f"dst{i}"]
        assert np.array_equal(oracle.resize_linear_exact_u8(g[f"src{i}"], dst.shape[1], dst.shape[0]), dst)
    assert np.array_equal(oracle.sep_gauss7_f32_u8(g["sep_src"]), g["sep_dst"])


def test_bird_front_end_live_cv2():
    cv2 = pytest.importorskip("cv2")
    was = cv2.useOptimized()
    cv2.setUseOptimized(False)          # OpenCV's own code paths (IPP changes getRectSubPix's float operation order)
    try:
        img, mask = cases.birdview_case(320, 4242, vehicle=(60, 100))
        orb = cv2.ORB_create(1500)
        det = orb.detect(img, mask)
        mine = oracle.bird_detect(img, mask, 1500)
        assert len(det) == len(mine) > 500
        for k, m in zip(det, mine):
            assert (k.pt[0], k.pt[1], k.octave) == (float(m["x"]), float(m["y"]), int(m["octave"]))
            assert np.float32(k.response) == m["response"] and np.float32(k.angle) == m["angle"] and np.float32(k.size) == m["size"]
        pts = np.array([k.pt for k in det], np.float32).reshape(-1, 1, 2)
        crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 40, 0.001)
        ref = cv2.cornerSubPix(img, pts.copy(), (5, 5), (-1, -1), crit).reshape(-1, 2)
        sub = oracle.corner_subpix(img, pts.reshape(-1, 2))
        assert np.array_equal(ref.view(np.uint32), sub.view(np.uint32))
        # points next to the image border take getRectSubPix's replicate path
        edge = np.array([[3.2, 4.1], [316.5, 200.2], [100.7, 317.9], [1.0, 318.0], [6.0, 6.0]], np.float32)
        ref = cv2.cornerSubPix(img, edge.reshape(-1, 1, 2).copy(), (5, 5), (-1, -1), crit).reshape(-1, 2)
        assert np.array_equal(ref.view(np.uint32), oracle.corner_subpix(img, edge).view(np.uint32))
        for k, p in zip(det, sub):
            k.pt = (float(p[0]), float(p[1]))
        kept, desc = orb.compute(img, det)
        moved = mine.copy()
        moved["x"], moved["y"] = sub[:, 0], sub[:, 1]
        k2, d2 = oracle.bird_compute(img, moved)
        assert len(kept) == len(k2) and np.array_equal(desc, d2)
        assert all((k.pt[0], k.pt[1]) == (float(m["x"]), float(m["y"])) for k, m in zip(kept, k2))
    finally:
        cv2.setUseOptimized(was)


def test_bird_detect_matches_python_transcription():
    """tests/ref_py/cv_orb_ref.py (cv2 primitives + transcribed glue incl. libstdc++ nth_element) == the C++ oracle."""
    pytest.importorskip("cv2")
    from ref_py import cv_orb_ref
    img, mask = cases.birdview_case(240, 515, vehicle=(40, 70))
    ref, _ = cv_orb_ref.detect(img, mask, 800)
    mine = oracle.bird_detect(img, mask, 800)
    assert len(ref) == len(mine) > 300
    for r, m in zip(ref, mine):
        assert (float(r[0]), float(r[1]), r[4]) == (float(m["x"]), float(m["y"]), int(m["octave"]))
        assert np.float32(r[2]) == m["response"] and np.float32(r[3]) == m["angle"]


def test_retain_best_order_with_ties():
    """FAST responses are small integers: the first retainBest works on heavily tied keys, where the element order out of
    std::nth_element is what decides which keypoints reach the Harris stage.  A low-texture image maximises ties."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8)
    img = (rng.integers(0, 2, (300, 300)) * 60 + 90).astype(np.uint8)
    img = cv2.resize(img[:60, :60], (300, 300), interpolation=cv2.INTER_NEAREST)      # 5x5 blocks: thousands of equal-score corners
    det = cv2.ORB_create(300).detect(img, None)
    mine = oracle.bird_detect(img, None, 300)
    assert len(det) == len(mine) > 100
    assert all((k.pt[0], k.pt[1], k.octave) == (float(m["x"]), float(m["y"]), int(m["octave"])) for k, m in zip(det, mine))
