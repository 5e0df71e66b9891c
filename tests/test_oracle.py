"""CPU tests: pin the C++ oracle against the committed golden vectors (cv2 4.13 outputs and the two
independent Python restatements), and against live cv2 where it is importable."""
import os

import numpy as np
import pytest

import cases
from helpers import GOLDEN, oracle, synth


def _npz(name):
    return np.load(os.path.join(GOLDEN, name))


# ---- OpenCV primitives ------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_primitives_vs_cv2_golden(tag):
    g = _npz("primitives.npz")
    img = g[f"img_{tag}"]
    want = g[f"resize_{tag}"]
    assert np.array_equal(oracle.resize_u8(img, want.shape[1], want.shape[0]), want)
    assert np.array_equal(oracle.gauss7_u8(img), g[f"gauss_{tag}"])
    for th in (7, 20):
        assert np.array_equal(oracle.fast9(img, th), g[f"fast{th}_{tag}"])
    cell = img[10:46, 20:57]          # strided view == cv::Mat ROI
    assert np.array_equal(oracle.fast9(cell, 7), g[f"fastcell_{tag}"])


def test_fast_atan2_golden():
    g = _npz("primitives.npz")
    got = np.array([oracle.fast_atan2(y, x) for y, x in g["atan_yx"]], np.float32)
    assert np.array_equal(got.view(np.uint32), g["atan_deg"].view(np.uint32))   # bit-exact


def test_cv_round_half_even():
    assert [oracle.cv_round(v) for v in (0.5, 1.5, 2.5, -0.5, -1.5, 2.4999, 2.5001)] == [0, 2, 2, 0, -2, 2, 3]


def test_fast_score_map_consistent_with_fast9():
    img = synth.synth_frame(80, 90, 5)
    sc = oracle.fast_score_map(img)
    for th in (5, 20):
        k = oracle.fast9(img, th, nms=False)
        m = np.zeros_like(sc, bool)
        m[k[:, 1], k[:, 0]] = True
        assert np.array_equal(m, sc >= th)
        assert np.array_equal(sc[k[:, 1], k[:, 0]], k[:, 2])


def test_primitives_vs_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for (h, w) in [(480, 752), (376, 1241), (41, 53), (7, 9)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        dw, dh = int(round(w / 1.2)), int(round(h / 1.2))
        assert np.array_equal(oracle.resize_u8(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))
        assert np.array_equal(oracle.gauss7_u8(img), cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
    img = synth.synth_frame(120, 160, 9)
    for th in (5, 7, 12, 20):
        kps = cv2.FastFeatureDetector_create(th, True).detect(img)
        want = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
        assert np.array_equal(oracle.fast9(img, th), want)


# ---- extractor --------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1_752x480", "bird_400x400", "small_320x240"])
def test_extractor_vs_python_cv2_golden(name):
    g = _npz(f"extract_{name}.npz")
    nf, ini, mn, seed = [int(v) for v in g["params"]]
    img = g["img"]
    assert np.array_equal(img, synth.synth_frame(img.shape[0], img.shape[1], seed)), "generator drifted"
    k, d = oracle.Extractor(nf, 1.2, 8, ini, mn)(img)
    assert len(k) == len(g["kps"])
    assert k.tobytes() == g["kps"].tobytes()
    assert np.array_equal(d, g["desc"])


def test_extractor_tables():
    ex = oracle.Extractor(1000, 1.2, 8, 20, 7)
    assert ex.features_per_level().tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert ex.umax().tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert oracle.Extractor(2000).features_per_level().tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert oracle.Extractor(4000).features_per_level().tolist() == [869, 724, 603, 503, 419, 349, 291, 242]
    assert np.array_equal(ex.scale_factors(), cases.SCALE_FACTORS)


def test_extractor_edge_cases():
    ex = oracle.Extractor(500, 1.2, 8, 20, 7)
    flat = np.full((240, 320), 128, np.uint8)
    k, d = ex(flat)
    assert len(k) == 0 and d.shape == (0, 32)
    # strided input (ROI of a larger buffer) == contiguous copy
    big = synth.synth_frame(300, 400, 21)
    roi = big[10:250, 30:350]
    k1, d1 = ex(roi)
    k2, d2 = ex(np.ascontiguousarray(roi))
    assert k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2)
    # upper levels too small for one FAST cell: defined as "no keypoints there"
    small = synth.synth_frame(100, 120, 22)
    k, _ = ex(small)
    assert len(k) > 0 and k["octave"].max() < 7
    # keypoints stay >= 19 px inside their level
    sf = ex.scale_factors()
    for lvl in range(8):
        kl = ex.level_keypoints(lvl)
        im = ex.level_image(lvl)
        if len(kl):
            assert kl["x"].min() >= 19 and kl["x"].max() < im.shape[1] - 19
            assert kl["y"].min() >= 19 and kl["y"].max() < im.shape[0] - 19
    assert sf[0] == 1.0


def test_octree_quota_and_determinism():
    img = synth.synth_frame(376, 1241, 2000)
    ex = oracle.Extractor(2000, 1.2, 8, 20, 7)
    k, d = ex(img)
    quota = ex.features_per_level()
    for lvl in range(8):
        n = int((k["octave"] == lvl).sum())
        assert quota[lvl] <= n <= quota[lvl] + 3, (lvl, n, quota[lvl])
    k2, d2 = ex(img)
    assert k.tobytes() == k2.tobytes() and np.array_equal(d, d2)
    # standalone octree entry == what the extractor produced, and is independent of candidate order
    # except through the "first maximum" rule, which keys on (cell row, cell col, y, x) order
    c = ex.level_candidates(0)
    im = ex.level_image(0)
    out = oracle.distribute_octree(c, 16, im.shape[1] - 16, 16, im.shape[0] - 16, int(quota[0]))
    kl = ex.level_keypoints(0)
    assert np.array_equal(out[:, 0] + 16, kl["x"].astype(np.int32)) and np.array_equal(out[:, 1] + 16, kl["y"].astype(np.int32))


# ---- matcher ----------------------------------------------------------------------------------------
def test_descriptor_distance_is_popcount():
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    want = np.unpackbits(a ^ b, axis=1).sum(1)
    got = [oracle.descriptor_distance(a[i], b[i]) for i in range(200)]
    assert got == want.tolist()
    assert oracle.descriptor_distance(a[0], a[0]) == 0
    assert oracle.descriptor_distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_knn2_vs_numpy():
    q = synth.synth_descriptors(64, 1)
    m = synth.synth_descriptors(500, 2)
    m[100] = q[3]
    m[200] = q[3]          # duplicate best: first index must win, second == best
    bi, bd, sd = oracle.hamming_knn2(q, m)
    D = np.unpackbits(q[:, None, :] ^ m[None, :, :], axis=2).sum(2)
    assert np.array_equal(bi, D.argmin(1))
    assert np.array_equal(bd, D.min(1))
    assert np.array_equal(sd, np.sort(D, 1)[:, 1])
    assert bi[3] == 100 and bd[3] == 0 and sd[3] == 0
    bi, bd, sd = oracle.hamming_knn2(q[:2], m[:1])
    assert sd.tolist() == [256, 256]
    bi, bd, sd = oracle.hamming_knn2(q[:2], m[:0])
    assert bi.tolist() == [-1, -1] and bd.tolist() == [256, 256]


def test_distinctive_descriptors_vs_python():
    """MapPoint[Bird]::ComputeDistinctiveDescriptors selection: oracle vs the literal transcription."""
    import matcher_py_ref as mref
    sizes = [0, 1, 2, 3, 4, 5, 8, 13, 32, 33, 40, 0, 7]
    desc, ptr = cases.distinctive_groups(sizes, 31)
    bi, bm = oracle.distinctive_descriptors(desc, ptr)
    for g, n in enumerate(sizes):
        want = mref.compute_distinctive_descriptors([desc[i] for i in range(ptr[g], ptr[g + 1])])
        assert (int(bi[g]), int(bm[g])) == want, f"group {g} (N={n})"
    # all descriptors equal: every median is 0, the first row wins
    same = np.tile(desc[:1], (6, 1))
    bi, bm = oracle.distinctive_descriptors(same, np.array([0, 6], np.int32))
    assert bi.tolist() == [0] and bm.tolist() == [0]


def _golden_frame():
    w, h = 620, 188
    kps, desc, uR, grid = cases.frame_case(500, w, h, 11, stereo_frac=0.4)
    F = oracle.Frame(kps, desc, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"], uR)
    q = cases.projection_queries(kps, desc, uR, w, h, 700, 12)
    blocked = (np.random.default_rng(13).random(len(kps)) < 0.1).astype(np.uint8)
    return F, q, blocked


def test_features_in_area_vs_python():
    import matcher_py_ref as mref
    kps, desc, uR, grid = cases.frame_case(300, 400, 400, 5)
    F = oracle.Frame(kps, desc, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"])
    P = mref.PyFrame(kps, desc, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"])
    rng = np.random.default_rng(6)
    for _ in range(300):
        x, y = rng.uniform(-30, 430, 2)
        r = float(rng.choice([4.0, 10.0, 15.0, 37.5]))
        lo, hi = [(-1, -1), (0, 0), (2, 3), (1, -1), (0, 4)][int(rng.integers(0, 5))]
        assert F.features_in_area(x, y, r, lo, hi).tolist() == P.features_in_area(x, y, r, lo, hi)


def test_matchers_vs_python_golden():
    g = _npz("matcher.npz")
    F, q, blocked = _golden_frame()
    for th in (1.0, 3.0):
        n, bi, bd, qk = oracle.search_by_projection(F, cases.SCALE_FACTORS, q["valid"], q["u"], q["v"], q["uR"], q["level"],
                                                    q["viewcos"], q["desc"], q["obs_pos"], blocked, th, 0.8)
        want = g[f"sbp_th{int(th)}"]
        assert n == want[0] and np.array_equal(qk, want[1:])
    for mode in (0, 1, 2):
        n, qk = oracle.search_by_projection_frame(F, cases.SCALE_FACTORS, q["valid"], q["u"], q["v"], q["invz"], q["level"],
                                                  q["angle"], q["desc"], q["obs_pos"], blocked, 7.0, 40.0, mode, True)
        want = g[f"sbpf_mode{mode}"]
        assert n == want[0] and np.array_equal(qk, want[1:])
    (k1, d1), (k2, d2), gr = cases.bird_pair(400, 200, 21)
    F2 = oracle.Frame(k2, d2, gr["min_x"], gr["min_y"], gr["inv_w"], gr["inv_h"])
    n, m12, _ = oracle.birdview_match(k1, d1, F2, None, 10, 0.99, True)
    assert n == g["bird_a"][0] and np.array_equal(m12, g["bird_a"][1:])
    prev = np.stack([k1["x"], k1["y"]], 1)
    n, m12, prev2 = oracle.birdview_match(k1, d1, F2, prev, 15, 0.99, True)
    assert n == g["bird_b"][0] and np.array_equal(m12, g["bird_b"][1:]) and np.array_equal(prev2, g["bird_b_prev"])
    has = (np.random.default_rng(22).random(len(k1)) < 0.6).astype(np.uint8)
    n, mk = oracle.search_by_match_bird_kf(k1, has, d1, F2, 15.0, 0.99, True)
    assert n == g["bird_kf"][0] and np.array_equal(mk, g["bird_kf"][1:])
    n, qk = oracle.search_by_projection_bird(F2, has, k1["x"] + 3, k1["y"] - 2, d1, None, None, 4.0, 0.99)
    assert n == g["bird_proj"][0] and np.array_equal(qk, g["bird_proj"][1:])
    t = cases.triangulation_case(300, 300, 620, 188, 31)
    for only_stereo in (0, 1):
        n, pairs = oracle.search_for_triangulation(t["k1"], t["d1"], t["uR1"], t["has1"], t["k2"], t["d2"], t["uR2"], t["has2"],
                                                   t["fv1"], t["fv2"], t["F12"], t["ex"], t["ey"], t["sf2"], t["sigma2"],
                                                   bool(only_stereo), True)
        want = g[f"tri_{only_stereo}"]
        assert n == want[0, 0] and np.array_equal(pairs, want[1:])


def test_matchers_vs_python_live_small():
    """Fresh seeds (not the committed ones) on small cases, both restatements run here."""
    import matcher_py_ref as mref
    for seed in (101, 202):
        kps, desc, uR, grid = cases.frame_case(200, 300, 200, seed, stereo_frac=0.5)
        F = oracle.Frame(kps, desc, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"], uR)
        P = mref.PyFrame(kps, desc, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"], uR)
        q = cases.projection_queries(kps, desc, uR, 300, 200, 300, seed + 1, jitter=2.0)
        n, bi, bd, qk = oracle.search_by_projection(F, cases.SCALE_FACTORS, q["valid"], q["u"], q["v"], q["uR"], q["level"],
                                                    q["viewcos"], q["desc"], q["obs_pos"], None, 2.0, 0.7)
        n2, qk2 = mref.search_by_projection(P, cases.SCALE_FACTORS, q["valid"], q["u"], q["v"], q["uR"], q["level"], q["viewcos"],
                                            q["desc"], q["obs_pos"], None, 2.0, 0.7)
        assert n == n2 and qk.tolist() == qk2
        (k1, d1), (k2, d2), gr = cases.bird_pair(150, 120, seed + 2, max_flips=10)
        F2 = oracle.Frame(k2, d2, gr["min_x"], gr["min_y"], gr["inv_w"], gr["inv_h"])
        P2 = mref.PyFrame(k2, d2, gr["min_x"], gr["min_y"], gr["inv_w"], gr["inv_h"])
        n, m12, _ = oracle.birdview_match(k1, d1, F2, None, 20, 0.9, True)
        n2, m122, _ = mref.birdview_match(k1, d1, P2, None, 20, 0.9, True)
        assert n == n2 and m12.tolist() == m122


def test_second_batch_matchers_vs_python_live():
    """SearchForInitialization, SearchByBoW x2 and the best-in-window family: oracle vs the literal Python
    transcriptions of the reference functions (fresh small cases)."""
    import matcher_py_ref as mref
    sf = cases.SCALE_FACTORS
    inv_s2 = (np.float32(1) / (sf * sf)).astype(np.float32)
    for seed in (301, 302):
        # SearchForInitialization on a front-camera grid
        (k1, d1), (k2, d2), gr = cases.bird_pair(260, 300, seed, shift=(4, 3), max_flips=12)
        F2 = oracle.Frame(k2, d2, gr["min_x"], gr["min_y"], gr["inv_w"], gr["inv_h"])
        P2 = mref.PyFrame(k2, d2, gr["min_x"], gr["min_y"], gr["inv_w"], gr["inv_h"])
        prev = np.stack([k1["x"], k1["y"]], 1)
        n, m12, p = oracle.search_for_initialization(k1, d1, F2, prev, 30, 0.9, True)
        n0, m120, p0 = mref.search_for_initialization(k1, d1, P2, prev, 30, 0.9, True)
        assert n == n0 and m12.tolist() == m120 and np.array_equal(p, p0)
        # SearchByBoW both variants
        b = cases.bow_case(220, 240, 400, 300, seed + 10)
        F2 = oracle.Frame(b["k2"], b["d2"], 0, 0, 64 / 400, 48 / 300)
        P2 = mref.PyFrame(b["k2"], b["d2"], 0, 0, 64 / 400, 48 / 300)
        for kf_kf in (False, True):
            n, out = oracle.search_by_bow(b["d1"], b["k1"]["angle"], b["valid1"], F2, b["valid2"], b["fv1"], b["fv2"], 0.75, True, kf_kf)
            n0, out0 = mref.search_by_bow(b["d1"], b["k1"]["angle"], b["valid1"], P2, b["valid2"], cases.csr_to_dict(b["fv1"]),
                                          cases.csr_to_dict(b["fv2"]), 0.75, True, kf_kf)
            assert n == n0 and n > 0 and out.tolist() == out0, (seed, kf_kf)
        # best-in-window family
        kps, desc, uR, grid = cases.frame_case(300, 400, 300, seed + 20, stereo_frac=0.5)
        F = oracle.Frame(kps, desc, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"], uR)
        P = mref.PyFrame(kps, desc, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"], uR)
        q = cases.best_window_queries(kps, desc, uR, 400, 300, 350, seed + 21, th=10.0)
        has = (np.random.default_rng(seed).random(len(kps)) < 0.2).astype(np.uint8)
        # SearchByProjection(Frame&, KeyFrame*, set, th, ORBdist): levels pred-1..pred+1, BLOCK|ORI
        n, bi, bd, qk = oracle.search_window_best(F, q["valid"], q["u"], q["v"], q["r"], q["pred"] - 1, q["pred"] + 1, q["desc"], None, q["angle"],
                                                  None, has, None, 64, oracle.WB_BLOCK | oracle.WB_ORI)
        n0, qk0 = mref.search_by_projection_kf(P, sf, q["valid"], q["u"], q["v"], q["pred"], q["angle"], q["desc"], has, 10.0, 64, True)
        assert n == n0 and qk.tolist() == qk0
        # SearchByProjection(KeyFrame*, Scw, ...): levels pred-1..pred, BLOCK, TH_LOW
        n, bi, bd, qk = oracle.search_window_best(F, q["valid"], q["u"], q["v"], q["r"], q["pred"] - 1, q["pred"], q["desc"], None, None,
                                                  None, has, None, 50, oracle.WB_BLOCK)
        n0, qk0 = mref.search_by_projection_scw(P, sf, q["valid"], q["u"], q["v"], q["pred"], q["desc"], has, 10.0)
        assert n == n0 and qk.tolist() == qk0
        # Fuse: independent, chi2 gate
        q3 = cases.best_window_queries(kps, desc, uR, 400, 300, 350, seed + 22, th=3.0)
        n, bi, bd, qk = oracle.search_window_best(F, q3["valid"], q3["u"], q3["v"], q3["r"], q3["pred"] - 1, q3["pred"], q3["desc"], q3["ur"],
                                                  None, None, None, inv_s2, 50, oracle.WB_CHI2)
        n0, bi0 = mref.fuse_best(P, sf, inv_s2, q3["valid"], q3["u"], q3["v"], q3["ur"], q3["pred"], q3["desc"], 3.0)
        assert n == n0 and n > 0 and bi.tolist() == bi0


def test_stereo_matches_vs_python_live():
    """Frame::ComputeStereoMatches: C++ oracle vs the literal Python transcription (tests/ref_py/stereo_py_ref.py)."""
    from stereo_py_ref import compute_stereo_matches as pyref
    for (h, w, nf, sh, seed) in [(240, 320, 500, -6, 77), (200, 400, 600, -25, 5)]:
        left = synth.synth_frame(h, w, seed)
        right = synth.shift_frame(left, sh, 0)
        right = np.clip(right.astype(int) + np.random.default_rng(seed).integers(-3, 4, right.shape), 0, 255).astype(np.uint8)
        eL, eR = oracle.Extractor(nf, 1.2, 8, 20, 7), oracle.Extractor(nf, 1.2, 8, 20, 7)
        kl, dl = eL(left)
        kr, dr = eR(right)
        n, ur, dp = oracle.compute_stereo_matches(eL, eR, kl, dl, kr, dr, 0.537, 386.1448)
        sf = eL.scale_factors()
        n0, ur0, dp0 = pyref([eL.level_image(l) for l in range(8)], [eR.level_image(l) for l in range(8)], kl, dl, kr, dr, sf,
                             (np.float32(1) / sf).astype(np.float32), 0.537, 386.1448)
        assert n == n0 and n > 100
        assert np.array_equal(ur.view(np.uint32), ur0.view(np.uint32)) and np.array_equal(dp.view(np.uint32), dp0.view(np.uint32))


def test_bow_transform_vs_python_live():
    """DBoW2 transform: C++ oracle vs the literal Python transcription (tests/ref_py/bow_py_ref.py)."""
    import bow_py_ref
    voc = cases.synthetic_vocabulary(10, 3, 1)
    V = oracle.Vocabulary(**voc)
    leaf = np.nonzero(voc["word_id"] >= 0)[0]
    feats = synth.perturb_descriptors(voc["node_desc"][np.random.default_rng(2).choice(leaf, 300)], 20, 3)
    for lu in (1, 2, 4):
        w, n, (bw, bv), (fn, fp, fi) = V.transform(feats, lu)
        w0, n0, v0, fv0 = bow_py_ref.transform(voc, feats, lu)
        assert w.tolist() == w0 and n.tolist() == n0
        assert bw.tolist() == [k for k, _ in v0] and bv.tolist() == [x for _, x in v0]
        assert fn.tolist() == [k for k, _ in fv0] and all(fi[fp[i]:fp[i + 1]].tolist() == fv0[i][1] for i in range(len(fv0)))


def test_is_in_frustum_matches_cv2_composition():
    """The oracle's restatement of the cv::Mat arithmetic in Frame::isInFrustum against cv2.gemm / cv2.norm."""
    import oracle
    from ref_py import frustum_py_ref
    k, d, _, _ = cases.frame_case(600, 1241, 376, 77)
    for seed in (1, 2, 3):
        pose, mp = cases.local_map_case(k, d, 1241, 376, 500, 900 + seed)
        P = oracle.camera_pose(**pose)
        n, iv, u, v, uR, lvl, vc = oracle.is_in_frustum(mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"], P, 0.5,
                                                        mp["candidate"])
        ref = frustum_py_ref.is_in_frustum(mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"], pose, 0.5, mp["candidate"])
        assert n == int(iv.sum()) and 0.2 * len(iv) < n < 0.9 * len(iv)          # the case exercises both outcomes
        assert np.array_equal(iv, ref["in_view"])
        for a, b in ((u, ref["u"]), (v, ref["v"]), (uR, ref["uR"]), (vc, ref["viewcos"])):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert np.array_equal(lvl, ref["level"])
        assert len(np.unique(lvl[iv == 1])) >= 4
