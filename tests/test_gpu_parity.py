"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs.  Bar: bit-exact keypoints (coordinates, octave, response, angle), descriptors, match indices and
Hamming distances.  Descriptor tolerance per BASELINE.json north_star: bit-exact wherever the angle agrees
within 1e-4 rad, mismatch rate <= 0.1 % (we expect and assert 0 on these inputs and report the rate)."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
from helpers import GOLDEN, KP_DTYPE, oracle, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import orb_slam_birdview_b200 as pkg
    pkg.load_library()
    return pkg


def _compare_extract(k_gpu, d_gpu, k_ref, d_ref, tag=""):
    assert len(k_gpu) == len(k_ref), f"{tag}: keypoint count {len(k_gpu)} vs {len(k_ref)}"
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        bad = np.nonzero(k_gpu[f] != k_ref[f])[0]
        assert len(bad) == 0, f"{tag}: field {f} differs at {bad[:5]}: {k_gpu[f][bad[:5]]} vs {k_ref[f][bad[:5]]}"
    dang = np.abs(k_gpu["angle"].astype(np.float64) - k_ref["angle"].astype(np.float64)) * np.pi / 180.0
    assert np.array_equal(k_gpu["angle"].view(np.uint32), k_ref["angle"].view(np.uint32)), f"{tag}: angles not bit-exact, max diff {dang.max()} rad"
    agree = dang <= 1e-4
    rows_bad = (d_gpu != d_ref).any(1)
    rate = float(rows_bad.mean()) if len(rows_bad) else 0.0
    assert not (rows_bad & agree).any() or rate <= 1e-3, f"{tag}: descriptor mismatch rate {rate}"
    assert rate == 0.0, f"{tag}: descriptor mismatch rate {rate} (budget 1e-3)"
    return rate


SHAPES = [
    ("c1", 480, 752, 1000, 20, 7, 1000),
    ("bird", 400, 400, 2000, 15, 5, 3001),
    ("kitti", 376, 1241, 2000, 20, 7, 2000),
    ("small", 240, 320, 500, 20, 7, 77),
    ("fisheye", 400, 950, 2000, 15, 5, 4000),
]


@pytest.mark.parametrize("name,h,w,nf,ini,mn,seed", SHAPES)
def test_extract_vs_oracle(pkg, name, h, w, nf, ini, mn, seed):
    img = synth.synth_frame(h, w, seed)
    ex = pkg.ORBextractor(nf, 1.2, 8, ini, mn, max_size=(w, h))
    k, d = ex(img)
    orc = oracle.Extractor(nf, 1.2, 8, ini, mn)
    k0, d0 = orc(img)
    # stage by stage first, so a failure names the stage
    pyr, blur = ex.image_pyramid(0, False), ex.image_pyramid(0, True)
    for lvl in range(8):
        assert np.array_equal(pyr[lvl], orc.level_image(lvl)), f"{name}: pyramid level {lvl}"
        assert np.array_equal(blur[lvl], orc.level_image(lvl, blurred=True)), f"{name}: blurred level {lvl}"
        c_gpu = ex.level_candidates(0, lvl)
        c_ref = orc.level_candidates(lvl)
        key = lambda c: np.lexsort((c[:, 2], c[:, 0], c[:, 1]))
        assert len(c_gpu) == len(c_ref), f"{name}: candidates level {lvl}: {len(c_gpu)} vs {len(c_ref)}"
        assert np.array_equal(c_gpu[key(c_gpu)], c_ref[key(c_ref)]), f"{name}: candidate set level {lvl}"
    _compare_extract(k, d, k0, d0, name)
    # idempotence on a reused context
    k2, d2 = ex(img)
    assert k.tobytes() == k2.tobytes() and np.array_equal(d, d2)


def test_small_host_calls_one_graph_paths(pkg):
    """One or two images per host call run as one graph whose first kernel reads the pinned staging rows and whose last kernel
    writes the pinned result mirror.  Rows of the caller's images need not be contiguous (a cv::Mat ROI) and widths need not be
    multiples of 16 (the staging pads them): strided views, two images per call, an odd width, the same context reused across
    shapes -- all byte-exact against the oracle, also on the replayed graph."""
    orc = oracle.Extractor(700, 1.2, 8, 20, 7)
    ex = pkg.ORBextractor(700, 1.2, 8, 20, 7, max_size=(640, 400), max_batch=2)
    for (h, w, seed) in ((300, 437, 31), (400, 640, 32), (300, 437, 33)):
        big = [synth.synth_frame(h + 10, w + 23, seed + 10 * i) for i in range(2)]
        views = [b[5:5 + h, 11:11 + w] for b in big]                  # stride = w + 23
        want = [orc(np.ascontiguousarray(v)) for v in views]
        for rep in range(3):                                          # warm-up, capture, replay
            ks, ds, ns = ex.extract_batch(views)
            for i, (k0, d0) in enumerate(want):
                _compare_extract(ks[i, :ns[i]], ds[i, :ns[i]], k0, d0, f"{w}x{h} strided pair, call {rep}")
            k, d = ex(views[1])
            _compare_extract(k, d, want[1][0], want[1][1], f"{w}x{h} strided single, call {rep}")


def test_replayed_extraction_graph_on_a_changing_sequence(pkg):
    """The replayed one-call graph (programmatic dependent launches, three branches, pinned mirror) on 24 different frames in a
    row: nothing of frame k-1 may leak into frame k (a kernel that started before its predecessor had finished would read the
    previous frame's pyramid, candidates or counters)."""
    h, w = 188, 620
    canvas = synth.synth_frame(h, w + 24 * 7 + 8, 4711)
    orc = oracle.Extractor(800, 1.2, 8, 20, 7)
    ex = pkg.ORBextractor(800, 1.2, 8, 20, 7, max_size=(w, h), max_batch=2)
    for i in range(24):
        img = np.ascontiguousarray(canvas[:, 7 * i:7 * i + w])
        if i % 5 == 4:
            img = np.ascontiguousarray(img[::-1])                     # a very different frame in between
        k0, d0 = orc(img)
        k, d = ex(img)
        _compare_extract(k, d, k0, d0, f"frame {i}")


@pytest.mark.parametrize("name", ["c1_752x480", "bird_400x400", "small_320x240"])
def test_extract_vs_committed_golden(pkg, name):
    g = np.load(os.path.join(GOLDEN, f"extract_{name}.npz"))
    nf, ini, mn, _ = [int(v) for v in g["params"]]
    img = g["img"]
    ex = pkg.ORBextractor(nf, 1.2, 8, ini, mn, max_size=(img.shape[1], img.shape[0]))
    k, d = ex(img)
    _compare_extract(k, d, g["kps"], g["desc"], name)


def test_extract_full_hd_4000(pkg):
    """BASELINE.json config 5 shape at full size against the oracle."""
    img = synth.synth_frame(1080, 1920, 5000)
    ex = pkg.ORBextractor(4000, 1.2, 8, 20, 7, max_size=(1920, 1080))
    k, d = ex(img)
    k0, d0 = oracle.Extractor(4000, 1.2, 8, 20, 7)(img)
    _compare_extract(k, d, k0, d0, "c5")


def test_extract_large_image_near_the_size_limit(pkg):
    """A 4000 x 3000 image (12 MP, the level-0 FAST region is close to the 4096-pixel coordinate limit of the packed candidates) with
    a 12000-feature budget: 133 x 99 FAST cells on level 0, ~3300 cell groups, the octree of level 0 with ~10^5 candidates."""
    img = synth.synth_frame(3000, 4000, 12000)
    ex = pkg.ORBextractor(12000, 1.2, 8, 20, 7, max_size=(4000, 3000))
    k, d = ex(img)
    k0, d0 = oracle.Extractor(12000, 1.2, 8, 20, 7)(img)
    assert len(k0) > 11000
    _compare_extract(k, d, k0, d0, "12 MP")


def test_extract_batch_independent_of_batching(pkg):
    imgs = [synth.synth_frame(376, 1241, 2000 + i) for i in range(5)]
    ex = pkg.ORBextractor(2000, 1.2, 8, 20, 7, max_size=(1241, 376), max_batch=5)
    K, D, N = ex.extract_batch(imgs)
    orc = oracle.Extractor(2000, 1.2, 8, 20, 7)
    for i, im in enumerate(imgs):
        k0, d0 = orc(im)
        _compare_extract(K[i, :N[i]], D[i, :N[i]], k0, d0, f"batch[{i}]")
    # one at a time on the same context, and in a different order: byte-identical
    for i in (3, 0):
        k, d = ex(imgs[i])
        assert k.tobytes() == K[i, :N[i]].tobytes() and np.array_equal(d, D[i, :N[i]])


def test_extract_edge_cases(pkg):
    ex = pkg.ORBextractor(500, 1.2, 8, 20, 7, max_size=(400, 300))
    k, d = ex(np.full((240, 320), 128, np.uint8))           # flat image: no keypoints, like the oracle
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = ex(np.empty((0, 0), np.uint8))                    # empty image: silent return
    assert len(k) == 0
    big = synth.synth_frame(300, 400, 21)
    roi = big[10:250, 30:350]                                # strided ROI == contiguous copy
    k1, d1 = ex(roi)
    k0, d0 = oracle.Extractor(500, 1.2, 8, 20, 7)(np.ascontiguousarray(roi))
    _compare_extract(k1, d1, k0, d0, "roi")
    small = synth.synth_frame(100, 120, 22)                  # top levels too small for one FAST cell
    k2, d2 = ex(small)
    k0, d0 = oracle.Extractor(500, 1.2, 8, 20, 7)(small)
    _compare_extract(k2, d2, k0, d0, "tiny levels")
    with pytest.raises(pkg.OrbB200Error):
        ex(synth.synth_frame(500, 700, 1))                   # larger than the context's max size
    with pytest.raises(pkg.OrbB200Error):
        ex(synth.synth_frame(340, 200, 2))                   # fewer pixels, but taller than max_h: per-row tables are sized by max_h


def test_getters_match_oracle(pkg):
    ex = pkg.ORBextractor(1000, 1.2, 8, 20, 7, max_size=(752, 480))
    orc = oracle.Extractor(1000, 1.2, 8, 20, 7)
    assert np.array_equal(ex.GetScaleFactors(), orc.scale_factors())
    assert np.array_equal(ex.ctx.features_per_level(), orc.features_per_level())
    assert ex.GetLevels() == 8
    sf = ex.GetScaleFactors()
    assert np.array_equal(ex.GetInverseScaleFactors(), (np.float32(1) / sf).astype(np.float32))
    assert np.array_equal(ex.GetScaleSigmaSquares(), (sf * sf).astype(np.float32))


# ---- Hamming -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("nq,nm", [(2000, 2000), (2000, 20000), (37, 1), (5, 0), (513, 129), (1, 70000)])
def test_knn2_vs_oracle(pkg, nq, nm):
    q = synth.synth_descriptors(nq, 100 + nq)
    m = synth.synth_descriptors(nm, 200 + nm)
    if nm > 10:
        rng = np.random.default_rng(5)
        plant = rng.integers(0, nm, max(1, nq // 10))
        m[plant] = synth.perturb_descriptors(q[rng.integers(0, nq, len(plant))], 30, 6)
        m[nm // 2] = q[0]
        m[nm // 2 + 3] = q[0]                                # exact duplicates: first index wins, second == best
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    bi, bd, sd = pkg.ORBmatcher(ctx).hamming_knn2(q, m)
    bi0, bd0, sd0 = oracle.hamming_knn2(q, m)
    assert np.array_equal(bi, bi0) and np.array_equal(bd, bd0) and np.array_equal(sd, sd0)


def test_distinctive_descriptors_vs_oracle(pkg):
    """MapPoint[Bird]::ComputeDistinctiveDescriptors selection, batched over landmarks (incl. empty groups, single
    observations, ties between duplicate observations, a group larger than one warp pass)."""
    ctx = pkg.Context(500, 1.2, 8, 20, 7, 320, 240, 1)
    m = pkg.ORBmatcher(ctx)
    sizes = [0, 1, 2, 3, 4, 5, 8, 13, 32, 33, 40, 0, 7, 129, 300] + [int(x) for x in np.random.default_rng(5).integers(1, 25, 400)]
    desc, ptr = cases.distinctive_groups(sizes, 32)
    bi, bm = m.ComputeDistinctiveDescriptors(desc, ptr)
    bi0, bm0 = oracle.distinctive_descriptors(desc, ptr)
    assert np.array_equal(bi, bi0) and np.array_equal(bm, bm0)
    # no landmarks / only empty landmarks
    bi, bm = m.ComputeDistinctiveDescriptors(np.zeros((0, 32), np.uint8), np.array([0], np.int32))
    assert len(bi) == 0
    bi, bm = m.ComputeDistinctiveDescriptors(np.zeros((0, 32), np.uint8), np.array([0, 0, 0], np.int32))
    assert bi.tolist() == [-1, -1] and bm.tolist() == [-1, -1]


def test_knn2_full_size_properties(pkg):
    """C4 at 2k x 200k: checked through size-independent properties + a sampled exact check."""
    nq, nm = 2000, 200000
    q = synth.synth_descriptors(nq, 1)
    m = synth.synth_descriptors(nm, 2)
    m[12345] = q[7]
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    bi, bd, sd = pkg.ORBmatcher(ctx).hamming_knn2(q, m)
    assert bi[7] == 12345 and bd[7] == 0
    assert (bd <= sd).all() and (bi >= 0).all() and (bi < nm).all()
    d = np.unpackbits(q ^ m[bi], axis=1).sum(1)
    assert np.array_equal(d, bd)                              # reported distance is the distance to the reported index
    sub = np.arange(0, nq, 97)
    bi0, bd0, sd0 = oracle.hamming_knn2(q[sub], m)
    assert np.array_equal(bi[sub], bi0) and np.array_equal(bd[sub], bd0) and np.array_equal(sd[sub], sd0)
    # permutation property: reversing the map order keeps best/second distances
    bi_r, bd_r, sd_r = pkg.ORBmatcher(ctx).hamming_knn2(q, m[::-1].copy())
    assert np.array_equal(bd_r, bd) and np.array_equal(sd_r, sd)


# ---- grid + windowed searches ---------------------------------------------------------------------------
def _frames(pkg, ctx, kps, desc, grid, uR=None):
    F = pkg.Frame(ctx, kps, desc, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"], uR)
    O = oracle.Frame(kps, desc, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"], uR)
    return F, O


def test_features_in_area(pkg):
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    kps, desc, uR, grid = cases.frame_case(1500, 1241, 376, 5)
    F, O = _frames(pkg, ctx, kps, desc, grid)
    rng = np.random.default_rng(6)
    for _ in range(60):
        x, y = rng.uniform(-40, 1280), rng.uniform(-40, 420)
        r = float(rng.choice([4.0, 10.0, 15.0, 37.5, 120.0]))
        lo, hi = [(-1, -1), (0, 0), (2, 3), (1, -1), (0, 4)][int(rng.integers(0, 5))]
        assert F.GetFeaturesInArea(x, y, r, lo, hi).tolist() == O.features_in_area(x, y, r, lo, hi).tolist()


@pytest.mark.parametrize("shift", [(-90.0, -35.0), (60.0, 25.0)])
def test_keypoints_outside_the_lookup_grid(pkg, shift):
    """Undistorted keypoints may fall outside the 64x48 grid (src/Frame.cc:549-559 PosInGrid returns false): they are in no
    cell, GetFeaturesInArea never returns them and the searches never match them (grid_build_kernel read
    unwritten cellItems beyond the in-grid count)."""
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    kps, desc, uR, grid = cases.frame_case(1800, 1241, 376, 21, stereo_frac=0.3)
    grid = dict(grid, min_x=np.float32(shift[0]), min_y=np.float32(shift[1]))      # mnMinX/mnMinY of a distorted camera
    if shift[0] > 0:    # a third of the keypoints left of / above the grid, some beyond its right/bottom edge
        kps["x"][::7] += np.float32(60.0)
    F, O = _frames(pkg, ctx, kps, desc, grid, uR)
    rng = np.random.default_rng(22)
    seen = set()
    for _ in range(40):
        x, y = rng.uniform(-60, 1300), rng.uniform(-60, 430)
        r = float(rng.choice([10.0, 37.5, 120.0, 400.0]))
        got = F.GetFeaturesInArea(x, y, r, -1, -1).tolist()
        assert got == O.features_in_area(x, y, r, -1, -1).tolist()
        seen.update(got)
    assert 0 < len(seen) < len(kps)
    q = cases.projection_queries(kps, desc, uR, 1241, 376, 2500, 23)
    m = pkg.ORBmatcher(ctx, 0.8)
    nm, bi, bd, qk = m.SearchByProjection(F, q["valid"], q["u"], q["v"], q["uR"], q["level"], q["viewcos"], q["desc"], q["obs_pos"], None, 3.0)
    nm0, bi0, bd0, qk0 = oracle.search_by_projection(O, cases.SCALE_FACTORS, q["valid"], q["u"], q["v"], q["uR"], q["level"], q["viewcos"],
                                                     q["desc"], q["obs_pos"], None, 3.0, 0.8)
    assert nm == nm0 and nm > 0 and np.array_equal(bi, bi0) and np.array_equal(qk, qk0)


@pytest.mark.parametrize("n,nq,w,h,seed,th,ratio", [(2000, 3000, 1241, 376, 11, 1.0, 0.8), (2000, 3000, 1241, 376, 12, 3.0, 0.8),
                                                     (500, 700, 620, 188, 13, 1.0, 0.8), (300, 2500, 200, 150, 14, 4.0, 0.9)])
def test_search_by_projection(pkg, n, nq, w, h, seed, th, ratio):
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    kps, desc, uR, grid = cases.frame_case(n, w, h, seed, stereo_frac=0.4)
    F, O = _frames(pkg, ctx, kps, desc, grid, uR)
    q = cases.projection_queries(kps, desc, uR, w, h, nq, seed + 1)
    blocked = (np.random.default_rng(seed).random(n) < 0.1).astype(np.uint8)
    m = pkg.ORBmatcher(ctx, ratio)
    nm, bi, bd, qk = m.SearchByProjection(F, q["valid"], q["u"], q["v"], q["uR"], q["level"], q["viewcos"], q["desc"], q["obs_pos"], blocked, th)
    nm0, bi0, bd0, qk0 = oracle.search_by_projection(O, cases.SCALE_FACTORS, q["valid"], q["u"], q["v"], q["uR"], q["level"], q["viewcos"],
                                                     q["desc"], q["obs_pos"], blocked, th, ratio)
    assert nm == nm0 and nm > 0
    assert np.array_equal(bi, bi0)
    assert np.array_equal(bd[bi >= 0], bd0[bi0 >= 0])
    assert np.array_equal(qk, qk0)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_search_by_projection_frame(pkg, mode):
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    w, h = 1241, 376
    kps, desc, uR, grid = cases.frame_case(2000, w, h, 31, stereo_frac=0.4)
    F, O = _frames(pkg, ctx, kps, desc, grid, uR)
    q = cases.projection_queries(kps, desc, uR, w, h, 2000, 32)
    for th in (7.0, 15.0, 30.0):
        nm, qk = pkg.ORBmatcher(ctx, 0.9, True).SearchByProjectionFrame(F, q["valid"], q["u"], q["v"], q["invz"], q["level"], q["angle"],
                                                                        q["desc"], q["obs_pos"], None, th, 40.0, mode)
        nm0, qk0 = oracle.search_by_projection_frame(O, cases.SCALE_FACTORS, q["valid"], q["u"], q["v"], q["invz"], q["level"], q["angle"],
                                                     q["desc"], q["obs_pos"], None, th, 40.0, mode, True)
        assert nm == nm0 and np.array_equal(qk, qk0), (mode, th)


@pytest.mark.parametrize("window,ratio,flips", [(10, 0.99, 25), (15, 0.99, 25), (20, 0.99, 25), (20, 0.7, 8)])
def test_birdview_match(pkg, window, ratio, flips):
    ctx = pkg.Context(2000, 1.2, 8, 15, 5, 64, 64)
    (k1, d1), (k2, d2), grid = cases.bird_pair(2000, 400, 41 + window, max_flips=flips)
    F2, O2 = _frames(pkg, ctx, k2, d2, grid)
    m = pkg.ORBmatcher(ctx, ratio, True)
    nm, m12, _ = m.BirdviewMatch(k1, d1, F2, window)
    nm0, m120, _ = oracle.birdview_match(k1, d1, O2, None, window, ratio, True)
    assert nm == nm0 and nm > 0 and np.array_equal(m12, m120)
    prev = np.stack([k1["x"] + 2, k1["y"] - 1], 1)
    nm, m12, p1 = m.BirdviewMatch(k1, d1, F2, window, prev)
    nm0, m120, p0 = oracle.birdview_match(k1, d1, O2, prev, window, ratio, True)
    assert nm == nm0 and np.array_equal(m12, m120) and np.array_equal(p1, p0)


def test_bird_kf_and_projection_bird(pkg):
    ctx = pkg.Context(2000, 1.2, 8, 15, 5, 64, 64)
    (k1, d1), (k2, d2), grid = cases.bird_pair(1500, 400, 51)
    F2, O2 = _frames(pkg, ctx, k2, d2, grid)
    has = (np.random.default_rng(52).random(len(k1)) < 0.6).astype(np.uint8)
    m = pkg.ORBmatcher(ctx, 0.99, True)
    for r in (15.0, 20.0):
        nm, out = m.SearchByMatchBird(k1, has, d1, F2, r)
        nm0, out0 = oracle.search_by_match_bird_kf(k1, has, d1, O2, r, 0.99, True)
        assert nm == nm0 and np.array_equal(out, out0)
    obs = (np.random.default_rng(53).random(len(k1)) < 0.8).astype(np.uint8)
    blocked = (np.random.default_rng(54).random(len(k2)) < 0.1).astype(np.uint8)
    nm, out = m.SearchByProjectionBird(F2, has, k1["x"] + 3, k1["y"] - 2, d1, obs, blocked, 4.0)
    nm0, out0 = oracle.search_by_projection_bird(O2, has, k1["x"] + 3, k1["y"] - 2, d1, obs, blocked, 4.0, 0.99)
    assert nm == nm0 and nm > 0 and np.array_equal(out, out0)


@pytest.mark.parametrize("only_stereo", [False, True])
def test_search_for_triangulation(pkg, only_stereo):
    ctx = pkg.Context(2000, 1.2, 8, 20, 7, 64, 64)
    t = cases.triangulation_case(2000, 2000, 1241, 376, 61, n_nodes=100)
    m = pkg.ORBmatcher(ctx, 0.6, True)
    n, pairs = m.SearchForTriangulation(t["k1"], t["d1"], t["uR1"], t["has1"], t["k2"], t["d2"], t["uR2"], t["has2"], t["fv1"], t["fv2"],
                                        t["F12"], t["ex"], t["ey"], t["sf2"], t["sigma2"], only_stereo)
    n0, pairs0 = oracle.search_for_triangulation(t["k1"], t["d1"], t["uR1"], t["has1"], t["k2"], t["d2"], t["uR2"], t["has2"], t["fv1"], t["fv2"],
                                                 t["F12"], t["ex"], t["ey"], t["sf2"], t["sigma2"], only_stereo, True)
    assert n == n0 and np.array_equal(pairs, pairs0)


def test_matchers_vs_committed_golden(pkg):
    """The Python-transcription golden vectors, through the CUDA path."""
    g = np.load(os.path.join(GOLDEN, "matcher.npz"))
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    w, h = 620, 188
    kps, desc, uR, grid = cases.frame_case(500, w, h, 11, stereo_frac=0.4)
    F = pkg.Frame(ctx, kps, desc, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"], uR)
    q = cases.projection_queries(kps, desc, uR, w, h, 700, 12)
    blocked = (np.random.default_rng(13).random(len(kps)) < 0.1).astype(np.uint8)
    for th in (1.0, 3.0):
        nm, bi, bd, qk = pkg.ORBmatcher(ctx, 0.8).SearchByProjection(F, q["valid"], q["u"], q["v"], q["uR"], q["level"], q["viewcos"],
                                                                     q["desc"], q["obs_pos"], blocked, th)
        want = g[f"sbp_th{int(th)}"]
        assert nm == want[0] and np.array_equal(qk, want[1:])
    (k1, d1), (k2, d2), gr = cases.bird_pair(400, 200, 21)
    F2 = pkg.Frame(ctx, k2, d2, gr["min_x"], gr["min_y"], gr["inv_w"], gr["inv_h"])
    nm, m12, _ = pkg.ORBmatcher(ctx, 0.99, True).BirdviewMatch(k1, d1, F2, 10)
    assert nm == g["bird_a"][0] and np.array_equal(m12, g["bird_a"][1:])


def test_stereo_step_device_matches_host_calls(pkg):
    """The batched C2 step with caller-made queries (orbb200_stereo_step_device / _host) == oracle."""
    ctx = pkg.Context(2000, 1.2, 8, 20, 7, 1241, 376, 6)
    L = ctx._L
    n, nq, w, h = 3, 600, 1241, 376
    imgs = np.stack([im for i in range(n) for im in (synth.synth_frame(h, w, 4242 + i), synth.shift_frame(synth.synth_frame(h, w, 4242 + i), -7, 0))])
    orc = oracle.Extractor(2000, 1.2, 8, 20, 7)
    ref = [orc(im) for im in imgs]
    qs = [synth.projection_queries(ref[2 * i][0], ref[2 * i][1], w, h, nq, 99 + i) for i in range(n)]
    q = {k: np.ascontiguousarray(np.stack([x[k] for x in qs])) for k in qs[0]}
    S = pkg.ProjQueries()
    S.q_valid, S.q_u, S.q_v, S.q_uR = (q[k].ctypes.data for k in ("valid", "u", "v", "uR"))
    S.q_level, S.q_viewcos, S.q_desc, S.q_obs_pos = (q[k].ctypes.data for k in ("level", "viewcos", "desc", "obs_pos"))
    cap = ctx.max_keypoints
    kps, desc, cnt = np.zeros((2 * n, cap), pkg.KP_DTYPE), np.zeros((2 * n, cap, 32), np.uint8), np.zeros(2 * n, np.int32)
    bi, bd, nm = np.zeros((n, nq), np.int32), np.zeros((n, nq), np.int32), np.zeros(n, np.int32)
    ctx.check(L.orbb200_stereo_step_host(ctx._h, imgs.ctypes.data, n, w, h, w, C.byref(S), nq, 1.0, 0.8, 0.0, 0.0, 64.0 / w, 48.0 / h,
                                         kps.ctypes.data, desc.ctypes.data, cap, cnt.ctypes.data, bi.ctypes.data, bd.ctypes.data, nm.ctypes.data), "stereo_step_host")
    ctx.sync()
    for i in range(n):
        kl, dl = ref[2 * i]
        F = oracle.Frame(kl, dl, np.float32(0), np.float32(0), np.float32(64.0 / w), np.float32(48.0 / h))
        n0, bi0, bd0, _ = oracle.search_by_projection(F, orc.scale_factors(), q["valid"][i], q["u"][i], q["v"][i], q["uR"][i], q["level"][i],
                                                      q["viewcos"][i], q["desc"][i], q["obs_pos"][i], None, 1.0, 0.8)
        assert kps[2 * i][:cnt[2 * i]].tobytes() == kl.tobytes() and np.array_equal(desc[2 * i][:len(kl)], dl)
        assert int(nm[i]) == n0 and n0 > 0 and np.array_equal(bi[i], bi0) and np.array_equal(bd[i][bi0 >= 0], bd0[bi0 >= 0])


def test_bench_steps_match_oracle(pkg):
    """What bench.py times (orbb200_frame_step_device / _host on both workloads, full-size frames) == oracle, frame by frame."""
    bench = pytest.importorskip("bench")
    res = bench.parity_check(n_frames=3)
    assert res["ok"], res


def test_cpp_shim(pkg, tmp_path):
    """The C++ ORB_SLAM2::ORBextractor shim (cpp/ORBextractor.h) called like Frame::ExtractORB == oracle."""
    import subprocess
    from helpers import ROOT
    drv = os.path.join(ROOT, "orb-slam-birdview_b200", "cpp", "shim_driver")
    if not os.path.exists(drv):
        subprocess.run(["make", "-s", "-C", os.path.dirname(drv)], check=True)
    img = synth.synth_frame(480, 752, 1000)
    raw, out = tmp_path / "in.raw", tmp_path / "out.bin"
    img.tofile(raw)
    subprocess.run([drv, str(raw), "752", "480", "1000", "20", "7", str(out)], check=True, timeout=120)
    buf = open(out, "rb").read()
    n = int(np.frombuffer(buf, np.int32, 1)[0])
    k = np.frombuffer(buf, pkg.KP_DTYPE, n, 4)
    d = np.frombuffer(buf, np.uint8, n * 32, 4 + 28 * n).reshape(n, 32)
    orc = oracle.Extractor(1000, 1.2, 8, 20, 7)
    k0, d0 = orc(img)
    _compare_extract(k, d, k0, d0, "cpp shim")
    off = 4 + 60 * n
    pw, ph = np.frombuffer(buf, np.int32, 2, off)
    lvl1 = np.frombuffer(buf, np.uint8, pw * ph, off + 8).reshape(ph, pw)
    assert np.array_equal(lvl1, orc.level_image(1))          # mvImagePyramid[1]
    # every level of mvImagePyramid (written into the pinned mirror by the extraction's own kernels), also for a width that is not a
    # multiple of 16 (the staged rows are padded) and after a second call on the same extractor
    for (h, w, nf, seed) in ((480, 752, 1000, 1000), (376, 1241, 2000, 5)):
        im = synth.synth_frame(h, w, seed)
        im.tofile(raw)
        pyr = tmp_path / "pyr.bin"
        subprocess.run([drv, str(raw), str(w), str(h), str(nf), "20", "7", str(out)], check=True, timeout=120,
                       env=dict(os.environ, ORBB200_SHIM_DUMP_PYRAMID=str(pyr)))
        pb = open(pyr, "rb").read()
        o2 = oracle.Extractor(nf, 1.2, 8, 20, 7)
        o2(im)
        nl, pos = int(np.frombuffer(pb, np.int32, 1)[0]), 4
        assert nl == 8
        for lvl in range(nl):
            lw, lh = (int(v) for v in np.frombuffer(pb, np.int32, 2, pos))
            got = np.frombuffer(pb, np.uint8, lw * lh, pos + 8).reshape(lh, lw)
            pos += 8 + lw * lh
            assert np.array_equal(got, o2.level_image(lvl)), (w, h, lvl)


def test_cpp_matcher_shim(pkg, tmp_path):
    """cpp/ORBmatcher_b200.cc -- the bodies of ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th),
    SearchForTriangulation(KF1, KF2, F12, ...), BirdviewMatch(const Frame&, const Frame&, ...) and
    SearchByMatchBird(Frame&, const Frame&, ...) over the C ABI -- driven through Frame / KeyFrame / MapPoint /
    MapPointBird objects like Tracking and LocalMapping do, against the oracle."""
    import subprocess
    from helpers import ROOT
    drv = os.path.join(ROOT, "orb-slam-birdview_b200", "cpp", "matcher_driver")
    if not os.path.exists(drv):
        subprocess.run(["make", "-s", "-C", os.path.dirname(drv)], check=True)
    rng = np.random.default_rng(77)
    w, h, nF, nq, th, ratio = 620, 188, 900, 1200, 1.0, 0.8
    kps, desc, uR, grid = cases.frame_case(nF, w, h, 71, stereo_frac=0.4)
    q = cases.projection_queries(kps, desc, uR, w, h, nq, 72)
    kp_obs = np.where(rng.random(nF) < 0.1, 3, np.where(rng.random(nF) < 0.1, 0, -1)).astype(np.int32)   # blocked / unobserved MapPoint / none
    bad = ((rng.random(nq) < 0.03) & (q["valid"] == 1)).astype(np.uint8)
    q_obs = np.where(q["obs_pos"] == 1, 2, 0).astype(np.int32)
    window, ratio_bird = 15, 0.99
    (k1, d1), (k2, d2), bgrid = cases.bird_pair(1500, 400, 73)
    hasmp1 = (rng.random(len(k1)) < 0.7).astype(np.uint8)
    case, out = tmp_path / "case.bin", tmp_path / "out.bin"
    with open(case, "wb") as f:
        np.array([nF, nq, len(k1), len(k2), window], np.int32).tofile(f)
        np.array([th, ratio, ratio_bird, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"], bgrid["inv_w"], bgrid["inv_h"]], np.float32).tofile(f)
        for a in (kps, desc, uR, kp_obs, q["valid"], bad, q["u"], q["v"], q["uR"], q["viewcos"], q["level"], q_obs, q["desc"],
                  k1, d1, hasmp1, k2, d2):
            np.ascontiguousarray(a).tofile(f)
        # two keyframes for SearchForTriangulation; pose chosen so that the epipole (:663-670) is exactly (cx, cy)
        t = cases.triangulation_case(1200, 1100, 1241, 376, 74, n_nodes=80)
        np.array([len(t["k1"]), len(t["k2"]), len(t["fv1"][0]), len(t["fv2"][0]), 0], np.int32).tofile(f)
        np.concatenate([[1, 1, t["ex"], t["ey"]], [0, 0, 0], [0, 0, 1], np.eye(3).ravel(), t["F12"].ravel()]).astype(np.float32).tofile(f)
        for a in (t["k1"], t["d1"], t["uR1"], t["has1"], t["k2"], t["d2"], t["uR2"], t["has2"], *t["fv1"], *t["fv2"], t["sf2"], t["sigma2"]):
            np.ascontiguousarray(a).tofile(f)
    subprocess.run([drv, str(case), str(out)], check=True, timeout=120)
    res = np.fromfile(out, np.int32)
    nm, mp_of_kp = res[0], res[1:1 + nF]
    o = 1 + nF
    nm_bird, m12 = res[o], res[o + 1:o + 1 + len(k1)]
    o += 1 + len(k1)
    nm_sbm, cur_mp = res[o], res[o + 1:o + 1 + len(k2)]
    dd = res[o + 1 + len(k2)]
    n_tri = res[o + 2 + len(k2)]
    tri_pairs = res[o + 3 + len(k2):o + 3 + len(k2) + 2 * n_tri].reshape(-1, 2)
    # SearchByProjection(Frame&, vector<MapPoint*>&, th) (src/ORBmatcher.cc:45-129)
    O = oracle.Frame(kps, desc, grid["min_x"], grid["min_y"], grid["inv_w"], grid["inv_h"], uR)
    valid = (q["valid"] == 1) & (bad == 0)
    nm0, bi0, bd0, qk0 = oracle.search_by_projection(O, cases.SCALE_FACTORS, valid.astype(np.uint8), q["u"], q["v"], q["uR"], q["level"],
                                                     q["viewcos"], q["desc"], (q_obs > 0).astype(np.uint8), (kp_obs > 0).astype(np.uint8), th, ratio)
    want = np.where(qk0 >= 0, qk0, np.where(kp_obs >= 0, -2, -1))
    assert nm == nm0 and nm > 0 and np.array_equal(mp_of_kp, want)
    # BirdviewMatch(const Frame&, const Frame&, vnMatches12, windowSize) (:1788-1899), SearchByMatchBird(Cur, Last, windowSize) (:1901-1921)
    O2 = oracle.Frame(k2, d2, bgrid["min_x"], bgrid["min_y"], bgrid["inv_w"], bgrid["inv_h"])
    nmb0, m120, _ = oracle.birdview_match(k1, d1, O2, None, window, ratio_bird, True)
    assert nm_bird == nmb0 and nmb0 > 0 and np.array_equal(m12, m120)
    want_cur = np.full(len(k2), -1, np.int32)
    carried = 0
    for k in range(len(k1)):
        if m120[k] >= 0 and hasmp1[k]:
            want_cur[m120[k]] = k
            carried += 1
    assert nm_sbm == carried and np.array_equal(cur_mp, want_cur)
    assert dd == int(np.unpackbits(d1[0] ^ d2[0]).sum())                     # DescriptorDistance (:1647-1663)
    # SearchForTriangulation(KF1, KF2, F12, vMatchedPairs, bOnlyStereo) (:657-823)
    n0, pairs0 = oracle.search_for_triangulation(t["k1"], t["d1"], t["uR1"], t["has1"], t["k2"], t["d2"], t["uR2"], t["has2"], t["fv1"], t["fv2"],
                                                 t["F12"], t["ex"], t["ey"], t["sf2"], t["sigma2"], False, True)
    assert n_tri == n0 and n0 > 0 and np.array_equal(tri_pairs, pairs0)


def test_cpp_birdview_shim(pkg, tmp_path):
    """cpp/BirdviewExtractor.h called like the birdview block of Frame::Frame (src/Frame.cc:328-342) == oracle."""
    import subprocess
    from helpers import ROOT
    drv = os.path.join(ROOT, "orb-slam-birdview_b200", "cpp", "shim_driver")
    subprocess.run(["make", "-s", "-C", os.path.dirname(drv)], check=True)
    img = synth.synth_frame(240, 320, 77)
    bimg, bmask = cases.birdview_case(400, 3101)
    raw, braw, mraw, out = tmp_path / "in.raw", tmp_path / "b.raw", tmp_path / "m.raw", tmp_path / "out.bin"
    img.tofile(raw)
    bimg.tofile(braw)
    bmask.tofile(mraw)
    subprocess.run([drv, str(raw), "320", "240", "500", "20", "7", str(out), str(braw), str(mraw), "400", "400"], check=True, timeout=120)
    buf = open(out, "rb").read()
    n = int(np.frombuffer(buf, np.int32, 1)[0])
    off = 4 + 60 * n
    pw, ph = np.frombuffer(buf, np.int32, 2, off)
    off += 8 + int(pw) * int(ph)
    nb = int(np.frombuffer(buf, np.int32, 1, off)[0])
    k = np.frombuffer(buf, pkg.KP_DTYPE, nb, off + 4)
    d = np.frombuffer(buf, np.uint8, nb * 32, off + 4 + 28 * nb).reshape(nb, 32)
    same = int(np.frombuffer(buf, np.int32, 1, off + 4 + 60 * nb)[0])
    k0, d0 = oracle.bird_extract(bimg, bmask, 2000)
    assert _same_kps(k, k0) and np.array_equal(d, d0) and same == 1


def test_search_for_initialization(pkg):
    ctx = pkg.Context(2000, 1.2, 8, 20, 7, 64, 64)
    (k1, d1), (k2, d2), grid = cases.bird_pair(2000, 600, 71, shift=(5, -4), max_flips=12)
    F2, O2 = _frames(pkg, ctx, k2, d2, grid)
    prev = np.stack([k1["x"], k1["y"]], 1)
    for win in (30, 100):
        nm, m12, p1 = pkg.ORBmatcher(ctx, 0.9, True).SearchForInitialization(k1, d1, F2, prev, win)
        nm0, m120, p0 = oracle.search_for_initialization(k1, d1, O2, prev, win, 0.9, True)
        assert nm == nm0 and nm > 0 and np.array_equal(m12, m120) and np.array_equal(p1, p0)


@pytest.mark.parametrize("kf_kf", [False, True])
def test_search_by_bow(pkg, kf_kf):
    ctx = pkg.Context(2000, 1.2, 8, 20, 7, 64, 64)
    b = cases.bow_case(2000, 2000, 1241, 376, 81, n_nodes=100)
    F2, O2 = _frames(pkg, ctx, b["k2"], b["d2"], b["grid"])
    for ratio in (0.75, 0.9):
        m = pkg.ORBmatcher(ctx, ratio, True)
        nm, out = m.SearchByBoW(b["d1"], b["k1"]["angle"], b["valid1"], F2, b["fv1"], b["fv2"], b["valid2"], kf_kf)
        nm0, out0 = oracle.search_by_bow(b["d1"], b["k1"]["angle"], b["valid1"], O2, b["valid2"], b["fv1"], b["fv2"], ratio, True, kf_kf)
        assert nm == nm0 and nm > 0 and np.array_equal(out, out0)
    # one big node (> the cached-candidate capacity): the overflow path
    n = 300
    k, d, _, grid = cases.frame_case(n, 400, 300, 82)
    d2 = synth.perturb_descriptors(d, 10, 83)
    fv = (np.array([7], np.int32), np.array([0, n], np.int32), np.arange(n, dtype=np.int32))
    F2, O2 = _frames(pkg, ctx, k, d2, grid)
    ones = np.ones(n, np.uint8)
    nm, out = pkg.ORBmatcher(ctx, 0.9, True).SearchByBoW(d, k["angle"], ones, F2, fv, fv, ones, kf_kf)
    nm0, out0 = oracle.search_by_bow(d, k["angle"], ones, O2, ones, fv, fv, 0.9, True, kf_kf)
    assert nm == nm0 and nm > 100 and np.array_equal(out, out0)


def test_search_window_best_family(pkg):
    """SearchByProjection(Frame,KF,set), SearchByProjection(KF,Scw), Fuse and SearchBySim3 configurations of the
    generic best-in-window entry point."""
    ctx = pkg.Context(2000, 1.2, 8, 20, 7, 64, 64)
    w, h = 1241, 376
    sf = cases.SCALE_FACTORS
    inv_s2 = (np.float32(1) / (sf * sf)).astype(np.float32)
    kps, desc, uR, grid = cases.frame_case(2000, w, h, 91, stereo_frac=0.5)
    F, O = _frames(pkg, ctx, kps, desc, grid, uR)
    has = (np.random.default_rng(92).random(len(kps)) < 0.2).astype(np.uint8)
    m = pkg.ORBmatcher(ctx, 0.9, True)
    WB = (pkg.ORBmatcher.WB_BLOCK, pkg.ORBmatcher.WB_URCHECK, pkg.ORBmatcher.WB_CHI2, pkg.ORBmatcher.WB_ORI)
    configs = [
        ("Frame,KF,set th=10 ORBdist=100", 10.0, -1, +1, WB[0] | WB[3], 100, False),
        ("Frame,KF,set th=3 ORBdist=64", 3.0, -1, +1, WB[0] | WB[3], 64, False),
        ("KF,Scw th=10", 10.0, -1, 0, WB[0], 50, False),
        ("Fuse th=3", 3.0, -1, 0, WB[2], 50, True),
        ("Sim3 th=7.5", 7.5, -1, 0, 0, 100, False),
        ("Cur,Last style with uRight check", 15.0, -1, +1, WB[0] | WB[1] | WB[3], 100, True),
    ]
    for name, th, lo, hi, flags, acc, use_aux in configs:
        q = cases.best_window_queries(kps, desc, uR, w, h, 2500, 93 + int(th), th=th)
        args = (q["valid"], q["u"], q["v"], q["r"], q["pred"] + lo, q["pred"] + hi, q["desc"], q["ur"] if use_aux else None,
                q["angle"] if flags & WB[3] else None, q["obs_pos"], has, inv_s2 if flags & WB[2] else None)
        nm, bi, bd, qk = m.search_window_best(F, *args, acc, flags)
        nm0, bi0, bd0, qk0 = oracle.search_window_best(O, *args, acc, flags)
        assert nm == nm0 and nm > 0, name
        assert np.array_equal(bi, bi0) and np.array_equal(bd[bi0 >= 0], bd0[bi0 >= 0]) and np.array_equal(qk, qk0), name


@pytest.mark.parametrize("h,w,nf,shift,seed", [(376, 1241, 2000, -7, 2000), (240, 320, 500, -6, 77), (200, 400, 600, -25, 5), (480, 752, 1000, -40, 9)])
def test_compute_stereo_matches(pkg, h, w, nf, shift, seed):
    """Frame::ComputeStereoMatches on the device-resident extraction results == oracle (bit-exact uRight, depth)."""
    left = synth.synth_frame(h, w, seed)
    right = synth.shift_frame(left, shift, 0)
    right = np.clip(right.astype(int) + np.random.default_rng(seed).integers(-3, 4, right.shape), 0, 255).astype(np.uint8)
    ex = pkg.ORBextractor(nf, 1.2, 8, 20, 7, max_size=(w, h), max_batch=2)
    K, D, N = ex.extract_batch([left, right])
    mb, mbf = 0.537, 386.1448
    nm, ur, dp = ex.ComputeStereoMatches(0, 1, mb, mbf)
    eL, eR = oracle.Extractor(nf, 1.2, 8, 20, 7), oracle.Extractor(nf, 1.2, 8, 20, 7)
    kl, dl = eL(left)
    kr, dr = eR(right)
    nm0, ur0, dp0 = oracle.compute_stereo_matches(eL, eR, kl, dl, kr, dr, mb, mbf)
    assert nm == nm0 and nm > 50
    assert np.array_equal(ur[:len(kl)].view(np.uint32), ur0.view(np.uint32))
    assert np.array_equal(dp[:len(kl)].view(np.uint32), dp0.view(np.uint32))
    # and the stereo frame feeds SearchByProjection's uRight consistency test
    import ctypes as C
    h_ = C.c_void_p()
    ctx = ex.ctx
    ctx.check(ctx._L.orbb200_frame_from_extract_stereo(ctx._h, C.byref(h_), 0, 0.0, 0.0, 64.0 / w, 48.0 / h))
    F = pkg.Frame.__new__(pkg.Frame)
    F.ctx, F._L, F._h, F.n = ctx, ctx._L, h_, ctx.max_keypoints
    O = oracle.Frame(kl, dl, 0.0, 0.0, 64.0 / w, 48.0 / h, ur0)
    q = cases.projection_queries(kl, dl, ur0, w, h, 1500, seed + 1)
    m = pkg.ORBmatcher(ctx, 0.8)
    nm1, bi, bd, qk = m.SearchByProjection(F, q["valid"], q["u"], q["v"], q["uR"], q["level"], q["viewcos"], q["desc"], q["obs_pos"], None, 2.0)
    nm2, bi0, bd0, qk0 = oracle.search_by_projection(O, cases.SCALE_FACTORS, q["valid"], q["u"], q["v"], q["uR"], q["level"], q["viewcos"],
                                                     q["desc"], q["obs_pos"], None, 2.0, 0.8)
    assert nm1 == nm2 and np.array_equal(bi, bi0) and np.array_equal(qk[:len(kl)], qk0)


@pytest.mark.parametrize("h,w,nf,scale,nlevels,ini,mn,seed", [
    (480, 640, 100, 1.2, 8, 20, 7, 1),        # tiny quota
    (480, 640, 8000, 1.2, 8, 20, 7, 2),       # quota far above what the image offers on upper levels
    (333, 517, 1500, 1.2, 8, 12, 5, 3),       # odd sizes (rows not 4-byte aligned), KITTI04-12 thresholds
    (400, 600, 1200, 1.5, 5, 20, 7, 4),       # other scale factor / level count
    (400, 600, 1200, 2.0, 4, 20, 7, 5),       # scale 2
    (376, 1241, 2000, 1.1, 12, 20, 7, 6),     # 12 levels
    (90, 1400, 800, 1.2, 3, 20, 7, 7),        # wide strip: many octree roots; 1-2 cell rows
    (107, 109, 300, 1.2, 2, 20, 7, 8),        # FAST cells up to 50 px wide
    (1080, 1920, 4000, 1.2, 8, 15, 5, 9),     # fisheye thresholds at full HD
])
def test_extract_parameter_sweep(pkg, h, w, nf, scale, nlevels, ini, mn, seed):
    img = synth.synth_frame(h, w, 8000 + seed)
    ex = pkg.ORBextractor(nf, scale, nlevels, ini, mn, max_size=(w, h))
    k, d = ex(img)
    orc = oracle.Extractor(nf, scale, nlevels, ini, mn)
    k0, d0 = orc(img)
    assert len(k0) > 0
    _compare_extract(k, d, k0, d0, f"sweep {h}x{w} nf={nf} s={scale} L={nlevels}")


def test_extract_noise_and_saturated_images(pkg):
    """Every pixel a corner candidate (uniform noise), saturated blocks, and a checkerboard."""
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, (240, 320), dtype=np.uint8)
    blocks = np.kron(rng.integers(0, 2, (15, 20)) * 255, np.ones((16, 16))).astype(np.uint8)
    checker = (np.indices((240, 320)).sum(0) // 7 % 2 * 255).astype(np.uint8)
    ex = pkg.ORBextractor(1000, 1.2, 8, 20, 7, max_size=(320, 240))
    orc = oracle.Extractor(1000, 1.2, 8, 20, 7)
    for name, img in (("noise", noise), ("blocks", blocks), ("checker", checker)):
        k, d = ex(img)
        k0, d0 = orc(img)
        _compare_extract(k, d, k0, d0, name)


@pytest.mark.parametrize("k,L,levelsup", [(10, 3, 1), (10, 3, 2), (10, 4, 4), (4, 6, 4)])
def test_bow_transform(pkg, k, L, levelsup):
    """DBoW2 transform (Frame::ComputeBoW) on a synthetic vocabulary tree: words, nodes, BowVector values (double bits)
    and FeatureVector == oracle; then the FeatureVectors drive SearchByBoW."""
    ctx = pkg.Context(2000, 1.2, 8, 20, 7, 64, 64)
    voc = cases.synthetic_vocabulary(k, L, seed=k * 10 + L)
    V = pkg.ORBVocabulary(ctx, **voc)
    O = oracle.Vocabulary(**voc)
    leaf = np.nonzero(voc["word_id"] >= 0)[0]
    rng = np.random.default_rng(5)
    feats = synth.perturb_descriptors(voc["node_desc"][rng.choice(leaf, 2000)], 40, 6)
    w, n, (bw, bv), (fn, fp, fi) = V.transform(feats, levelsup)
    w0, n0, (bw0, bv0), (fn0, fp0, fi0) = O.transform(feats, levelsup)
    assert np.array_equal(w[:len(w0)], w0) and np.array_equal(n[:len(n0)], n0)
    assert np.array_equal(bw, bw0) and np.array_equal(bv.view(np.uint64), bv0.view(np.uint64))
    assert np.array_equal(fn, fn0) and np.array_equal(fp, fp0) and np.array_equal(fi, fi0)
    assert abs(bv.sum() - 1.0) < 1e-9 and len(bw) > 50


def test_extract_bow_searchbybow_chain(pkg):
    """extract two views -> BoW transform of the device-resident descriptors -> SearchByBoW(KF, F): the chain of
    Tracking::TrackReferenceKeyFrame (src/Tracking.cc:1022-1035), each step against the oracle."""
    h, w = 376, 1241
    a = synth.synth_frame(h, w, 2000)
    b = synth.shift_frame(a, 5, 2)
    ex = pkg.ORBextractor(2000, 1.2, 8, 20, 7, max_size=(w, h), max_batch=2)
    K, D, N = ex.extract_batch([a, b])
    voc = cases.synthetic_vocabulary(10, 3, seed=9)
    V = pkg.ORBVocabulary(ex.ctx, **voc)
    O = oracle.Vocabulary(**voc)
    fvs = []
    for i in range(2):
        _, _, (bw, bv), fv = V.transform(None, 2, img_index=i)
        _, _, (bw0, bv0), fv0 = O.transform(D[i, :N[i]], 2)
        assert np.array_equal(bw, bw0) and np.array_equal(bv.view(np.uint64), bv0.view(np.uint64))
        assert all(np.array_equal(x, y) for x, y in zip(fv, fv0))
        fvs.append(fv)
    grid = (0.0, 0.0, 64.0 / w, 48.0 / h)
    F2 = pkg.Frame(ex.ctx, K[1, :N[1]], D[1, :N[1]], *grid)
    O2 = oracle.Frame(K[1, :N[1]], D[1, :N[1]], *grid)
    valid1 = (np.random.default_rng(3).random(N[0]) < 0.8).astype(np.uint8)
    nm, out = pkg.ORBmatcher(ex.ctx, 0.7, True).SearchByBoW(D[0, :N[0]], K[0, :N[0]]["angle"], valid1, F2, fvs[0], fvs[1])
    nm0, out0 = oracle.search_by_bow(D[0, :N[0]], K[0, :N[0]]["angle"], valid1, O2, None, fvs[0], fvs[1], 0.7, True, False)
    assert nm == nm0 and nm > 100 and np.array_equal(out, out0)


@pytest.mark.parametrize("n,seed", [(3000, 61), (700, 62), (1, 63)])
def test_is_in_frustum(pkg, n, seed):
    """Frame::isInFrustum + PredictScale over a device-resident map: flags and levels exact, floats bitwise."""
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    w, h = 1241, 376
    kps, desc, _, _ = cases.frame_case(2000, w, h, seed)
    pose, mp = cases.local_map_case(kps, desc, w, h, n, seed + 100)
    M = pkg.LocalMap(ctx, mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"], mp["desc"])
    for cand, lim in ((mp["candidate"], 0.5), (None, 0.5), (mp["candidate"], 0.9)):
        k, iv, u, v, uR, lvl, vc = M.isInFrustum(pkg.CameraPose.make(**pose), lim, cand)
        k0, iv0, u0, v0, uR0, lvl0, vc0 = oracle.is_in_frustum(mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"],
                                                               oracle.camera_pose(**pose), lim, cand)
        assert k == k0 and np.array_equal(iv, iv0) and np.array_equal(lvl, lvl0)
        for a, b in ((u, u0), (v, v0), (uR, uR0), (vc, vc0)):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    if n > 100:
        assert 0.2 * n < k0 < 0.9 * n


def test_search_local_points(pkg):
    """Tracking::SearchLocalPoints: projection on the device feeding SearchByProjection == oracle isInFrustum + oracle
    SearchByProjection."""
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    w, h = 1241, 376
    kps, desc, uR, grid = cases.frame_case(2000, w, h, 71, stereo_frac=0.4)
    F, O = _frames(pkg, ctx, kps, desc, grid, uR)
    pose, mp = cases.local_map_case(kps, desc, w, h, 3000, 72)
    blocked = (np.random.default_rng(73).random(len(kps)) < 0.1).astype(np.uint8)
    M = pkg.LocalMap(ctx, mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"], mp["desc"])
    for th in (1.0, 3.0, 5.0):
        k, iv, (u, v, uRq, lvl, vc), nm, bi, bd, qk = pkg.ORBmatcher(ctx, 0.8).SearchLocalPoints(
            F, M, pkg.CameraPose.make(**pose), mp["candidate"], mp["obs_pos"], blocked, th)
        k0, iv0, u0, v0, uR0, lvl0, vc0 = oracle.is_in_frustum(mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"],
                                                               oracle.camera_pose(**pose), 0.5, mp["candidate"])
        nm0, bi0, bd0, qk0 = oracle.search_by_projection(O, cases.SCALE_FACTORS, iv0, u0, v0, uR0, lvl0, vc0, mp["desc"], mp["obs_pos"],
                                                         blocked, th, 0.8)
        assert k == k0 and np.array_equal(iv, iv0) and np.array_equal(lvl, lvl0)
        assert nm == nm0 and nm > 50, (th, nm, nm0)
        assert np.array_equal(bi, bi0) and np.array_equal(bd[bi >= 0], bd0[bi0 >= 0]) and np.array_equal(qk, qk0)


# ---- birdview front-end: cv::ORB detect + cornerSubPix + compute (src/Frame.cc:328-342) -------------------------------
def _same_kps(a, b):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in a.dtype.names)


@pytest.mark.parametrize("size,seed,with_mask,nf", [(400, 3101, True, 2000), (384, 3102, False, 2000), ((500, 360), 3103, True, 2000),
                                                    (240, 99, True, 700), ((131, 97), 7, False, 300)])
def test_bird_detect_subpix_compute(pkg, size, seed, with_mask, nf):
    """Every stage against the oracle (itself pinned bit-exactly on cv2 4.13): keypoints incl. ORDER, Harris responses,
    angles; refined corners bitwise; descriptors exact."""
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    img, mask = cases.birdview_case(size, seed)
    mask = mask if with_mask else None
    B = pkg.BirdviewORB(ctx, nf)
    det = B.detect(img, mask)
    det0 = oracle.bird_detect(img, mask, nf)
    assert _same_kps(det, det0) and len(det0) > 50
    pts = np.stack([det0["x"], det0["y"]], 1)
    sub = B.cornerSubPix(img, pts)
    sub0 = oracle.corner_subpix(img, pts)
    assert np.array_equal(sub.view(np.uint32), sub0.view(np.uint32))
    moved = det0.copy()
    moved["x"], moved["y"] = sub0[:, 0], sub0[:, 1]
    k, d = B.compute(img, moved)
    k0, d0 = oracle.bird_compute(img, moved)
    assert _same_kps(k, k0) and np.array_equal(d, d0)
    k, d = B(img, mask)                                        # the fused device pipeline
    k0, d0 = oracle.bird_extract(img, mask, nf)
    assert _same_kps(k, k0) and np.array_equal(d, d0)


@pytest.mark.parametrize("name", ["400", "384_nomask", "500x360"])
def test_bird_extract_golden(pkg, name):
    """The device pipeline against vectors produced by cv2 itself (tests/golden/make_golden_bird.py)."""
    g = np.load(os.path.join(GOLDEN, f"bird_orb_{name}.npz"))
    w, h = (int(v) for v in g["size"])
    img, mask = cases.birdview_case((w, h), int(g["seed"]))
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    k, d = pkg.BirdviewORB(ctx, 2000)(img, mask if int(g["with_mask"]) else None)
    assert _same_kps(k, g["kps"]) and np.array_equal(d, g["desc"])


def test_bird_corner_subpix_border_and_windows(pkg):
    """Corners whose window leaves the image (getRectSubPix's replicate path), other window sizes, unsorted octaves in
    compute, and a batch call."""
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    img, mask = cases.birdview_case(320, 4242, vehicle=(60, 100))
    B = pkg.BirdviewORB(ctx, 1500)
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform(0, 319, (300, 2)), [[3.2, 4.1], [316.5, 200.2], [100.7, 317.9], [1.0, 318.0], [6.0, 6.0], [0, 0], [319, 319]]]).astype(np.float32)
    for win, it, eps in (((5, 5), 40, 0.001), ((3, 3), 10, 0.01), ((7, 4), 100, 0.0), ((1, 1), 5, 0.1)):
        a = B.cornerSubPix(img, pts, win, it, eps)
        b = oracle.corner_subpix(img, pts, win, it, eps)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), win
    det = oracle.bird_detect(img, mask, 1500)
    perm = rng.permutation(len(det))
    k, d = B.compute(img, det[perm])
    k0, d0 = oracle.bird_compute(img, det[perm])
    assert _same_kps(k, k0) and np.array_equal(d, d0)
    imgs, masks = zip(*[cases.birdview_case(320, 600 + i, vehicle=(60, 100)) for i in range(3)])
    ks, ds = B.extract_batch(list(imgs), list(masks))
    for i in range(3):
        k0, d0 = oracle.bird_extract(imgs[i], masks[i], 1500)
        assert _same_kps(ks[i], k0) and np.array_equal(ds[i], d0)
    # batches of >= 8 images refine the corners with the one-thread-per-corner kernel
    imgs, masks = zip(*[cases.birdview_case(200, 700 + i, vehicle=(40, 60)) for i in range(9)])
    ks, ds = pkg.BirdviewORB(ctx, 600).extract_batch(list(imgs), list(masks))
    for i in range(9):
        k0, d0 = oracle.bird_extract(imgs[i], masks[i], 600)
        assert _same_kps(ks[i], k0) and np.array_equal(ds[i], d0)


def test_bird_retain_best_ties(pkg):
    """Heavily tied FAST scores: the survivors and their order come from the literal std::nth_element sequence."""
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    rng = np.random.default_rng(8)
    blocks = (rng.integers(0, 2, (60, 60)) * 60 + 90).astype(np.uint8)
    img = np.repeat(np.repeat(blocks, 5, 0), 5, 1)
    a = pkg.BirdviewORB(ctx, 300).detect(img, None)
    b = oracle.bird_detect(img, None, 300)
    assert _same_kps(a, b) and len(b) > 100


def test_bird_retain_best_dense_and_tall(pkg):
    """The row-major ordering of a level's corners has two forms (counting sort by row when the row cursors fit the level's
    shared-memory carve, bitonic network otherwise) and the selection three shared-memory tiers in batches plus the one-launch form
    for one or two images: a noise image whose level 0 holds several thousand corners (largest tier) and a tall, sparse one (more
    rows than the smallest tier has cursor slots), each alone and in a batch of three."""
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    rng = np.random.default_rng(12)
    dense = np.repeat(np.repeat(rng.integers(0, 256, (150, 150)).astype(np.uint8), 2, 0), 2, 1)          # 300 x 300, 2 x 2 noise blocks
    tall = np.full((2304, 128), 100, np.uint8)                                                             # a narrow noise strip: ~800 corners on 2304 rows
    tall[:, 60:68] = np.repeat(np.repeat(rng.integers(0, 256, (1152, 4)).astype(np.uint8), 2, 0), 2, 1)
    for img, nf in ((dense, 2000), (tall, 400)):
        want_k, want_d = oracle.bird_extract(img, None, nf)
        assert len(want_k) > 100
        B = pkg.BirdviewORB(ctx, nf)
        k, d = B(img, None)
        assert _same_kps(k, want_k) and np.array_equal(d, want_d)
        ks, ds = B.extract_batch([img, img.copy(), img.copy()], None)
        for i in range(3):
            assert _same_kps(ks[i], want_k) and np.array_equal(ds[i], want_d), i
    n0 = len(oracle.bird_detect(dense, None, 100000))
    assert n0 > 5120, n0                                                                                  # the dense case really is in the last tier


def test_bird_single_call_mask_cache(pkg):
    """One image per call keeps the last mask's pyramid on the device and skips its upload when the next call brings the same mask:
    same mask, changed mask, no mask, a detect() call in between (it rewrites the mask slot), a strided mask -- each equal to the oracle."""
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    B = pkg.BirdviewORB(ctx, 800)
    imgs = [cases.birdview_case(240, 40 + i)[0] for i in range(6)]
    mask_a = cases.birdview_case(240, 40)[1]
    mask_b = mask_a.copy()
    mask_b[:60, :] = 0
    wide = np.zeros((240, 300), np.uint8)
    wide[:, :240] = mask_b
    seq = [(0, mask_a), (1, mask_a), (2, mask_b), (3, mask_b), (4, None), (5, mask_b), (0, wide[:, :240]), (1, mask_a)]
    for step, (i, m) in enumerate(seq):
        k, d = B(imgs[i], m)
        k0, d0 = oracle.bird_extract(imgs[i], None if m is None else np.ascontiguousarray(m), 800)
        assert _same_kps(k, k0) and np.array_equal(d, d0), step
        if step == 4:
            det = B.detect(imgs[2], mask_a)                  # another entry point writes image 0's mask slot
            assert _same_kps(det, oracle.bird_detect(imgs[2], mask_a, 800))


def test_bird_edge_cases(pkg):
    """No corners, everything masked, images too small for the 31-pixel edge threshold, strided input, empty inputs."""
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    B = pkg.BirdviewORB(ctx, 500)
    flat = np.full((200, 200), 128, np.uint8)
    k, d = B(flat, None)
    assert len(k) == 0 and d.shape == (0, 32)
    img, mask = cases.birdview_case(200, 31)
    k, d = B(img, np.zeros_like(mask))
    assert len(k) == 0
    small = synth.synth_frame(60, 60, 5)                       # <= 2 * edgeThreshold: runByImageBorder clears every level
    assert len(B.detect(small, None)) == 0 and len(oracle.bird_detect(small, None, 500)) == 0
    med = synth.synth_frame(90, 110, 6)                        # only the first levels are wider than 62 pixels
    a, b = B.detect(med, None), oracle.bird_detect(med, None, 500)
    assert _same_kps(a, b) and len(b) > 0
    k, d = B(med, None)
    k0, d0 = oracle.bird_extract(med, None, 500)
    assert _same_kps(k, k0) and np.array_equal(d, d0)
    big = np.zeros((200, 260), np.uint8)                       # row stride != width
    big[:, :200] = img
    k, d = B(big[:, :200], mask)
    k0, d0 = oracle.bird_extract(img, mask, 500)
    assert _same_kps(k, k0) and np.array_equal(d, d0)
    assert len(B.cornerSubPix(img, np.zeros((0, 2), np.float32))) == 0
    k, d = B.compute(img, np.zeros(0, KP_DTYPE))
    assert len(k) == 0
    with pytest.raises(pkg.OrbB200Error):
        B.cornerSubPix(img, [[50, 50]], (9, 9))                # window half-size beyond the supported 7


def test_frustum_empty_and_degenerate(pkg):
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    kps, desc, _, _ = cases.frame_case(500, 640, 480, 5)
    pose, mp = cases.local_map_case(kps, desc, 640, 480, 64, 6)
    M = pkg.LocalMap(ctx, mp["pos"][:0], mp["normal"][:0], mp["max_distance"][:0], mp["min_distance"][:0], mp["desc"][:0])
    k, iv, *_ = M.isInFrustum(pkg.CameraPose.make(**pose))
    assert k == 0 and len(iv) == 0
    # points at the camera centre, on the image border, zero distance bounds
    pos = mp["pos"].copy()
    pos[0] = pose["Ow"]
    mx, mn = mp["max_distance"].copy(), mp["min_distance"].copy()
    mx[1], mn[1] = 0.0, 0.0
    M = pkg.LocalMap(ctx, pos, mp["normal"], mx, mn, mp["desc"])
    k, iv, u, v, uR, lvl, vc = M.isInFrustum(pkg.CameraPose.make(**pose))
    k0, iv0, u0, v0, uR0, lvl0, vc0 = oracle.is_in_frustum(pos, mp["normal"], mx, mn, oracle.camera_pose(**pose))
    assert k == k0 and np.array_equal(iv, iv0) and np.array_equal(lvl, lvl0)
    assert np.array_equal(u.view(np.uint32), u0.view(np.uint32)) and np.array_equal(vc.view(np.uint32), vc0.view(np.uint32))


# ---- the batched north-star frame step (stereo front camera + birdview), device-resident -----------------------------------
def _oracle_frame_step(seq, mp, frames, nfeat, bird_nf, w, h, bw, bh, th, ratio, window, bird_ratio, prev_bird=None):
    """The same frames through the oracle, one by one: what Frame::Frame + SearchLocalPoints + SearchByMatchBird compute."""
    ex, ex2 = oracle.Extractor(nfeat, 1.2, 8, 20, 7), oracle.Extractor(nfeat, 1.2, 8, 20, 7)
    res = []
    for i in frames:
        kl, dl = ex(seq["imgs"][2 * i])
        kr, dr = ex2(seq["imgs"][2 * i + 1])
        nst, ur, dep = oracle.compute_stereo_matches(ex, ex2, kl, dl, kr, dr, 0.537, 386.1448)
        O = oracle.Frame(kl, dl, np.float32(0), np.float32(0), np.float32(64.0 / w), np.float32(48.0 / h), ur)
        k0, iv, u, v, uR, lvl, vc = oracle.is_in_frustum(mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"],
                                                         oracle.camera_pose(**seq["poses"][i]), 0.5, None)
        nm, bi, bd, qk = oracle.search_by_projection(O, ex.scale_factors(), iv, u, v, uR, lvl, vc, mp["desc"], None, None, th, ratio)
        bk, bdsc = oracle.bird_extract(seq["bird_imgs"][i], seq["bird_mask"], bird_nf)
        m12, nbm = None, 0
        if prev_bird is not None:
            OB = oracle.Frame(bk, bdsc, np.float32(0), np.float32(0), np.float32(64.0 / bw), np.float32(48.0 / bh))
            nbm, m12, _ = oracle.birdview_match(prev_bird[0], prev_bird[1], OB, None, window, bird_ratio, True)
        prev_bird = (bk, bdsc)
        res.append(dict(kl=kl, dl=dl, kr=kr, dr=dr, ur=ur, dep=dep, nm=nm, bi=bi, bd=bd, bk=bk, bdsc=bdsc, m12=m12, nbm=nbm))
    return res, prev_bird


def test_frame_step_vs_oracle(pkg):
    """orbb200_frame_step_host: two consecutive calls (the second chained to the first) against the oracle frame by frame:
    keypoints + descriptors of both cameras, mvuRight / mvDepth bits, SearchLocalPoints matches and distances, birdview
    keypoints + descriptors, BirdviewMatch(previous, current) incl. across the call boundary."""
    w, h, bw, bh, nfeat, bnf, nmap = 620, 188, 200, 200, 1000, 600, 1500
    seq = synth.northstar_sequence(7, 11, w=w, h=h, bird=(bw, bh), vehicle=(40, 60))
    orc = oracle.Extractor(nfeat, 1.2, 8, 20, 7)
    mp = synth.northstar_map(seq, lambda im: orc(im), nmap, 3)
    ctx = pkg.Context(nfeat, 1.2, 8, 20, 7, w, h, 8)
    M = pkg.LocalMap(ctx, mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"], mp["desc"])
    step = pkg.FrameStep(ctx, w, h, M, mb=0.537, mbf=386.1448, th=1.0, nnratio=0.8, bird_size=(bw, bh), bird_nfeatures=bnf,
                         bird_mask=seq["bird_mask"], bird_window=15, bird_nnratio=0.99)
    poses = [pkg.CameraPose.make(**p) for p in seq["poses"]]
    prev = None
    for (a, b, chain) in ((0, 4, False), (4, 7, True)):
        out = step(seq["imgs"][2 * a:2 * b], seq["bird_imgs"][a:b], poses[a:b], chain=chain)
        want, prev = _oracle_frame_step(seq, mp, range(a, b), nfeat, bnf, w, h, bw, bh, 1.0, 0.8, 15, 0.99, prev)
        for j, r in enumerate(want):
            nl, nr = out["counts"][2 * j], out["counts"][2 * j + 1]
            assert out["kps"][2 * j][:nl].tobytes() == r["kl"].tobytes() and out["kps"][2 * j + 1][:nr].tobytes() == r["kr"].tobytes()
            assert np.array_equal(out["desc"][2 * j][:nl], r["dl"]) and np.array_equal(out["desc"][2 * j + 1][:nr], r["dr"])
            assert np.array_equal(out["u_right"][j][:nl].view(np.uint32), r["ur"].view(np.uint32))
            assert np.array_equal(out["depth"][j][:nl].view(np.uint32), r["dep"].view(np.uint32))
            assert out["map_nmatches"][j] == r["nm"] and r["nm"] > 100
            assert np.array_equal(out["map_best_idx"][j], r["bi"])
            assert np.array_equal(out["map_best_dist"][j][r["bi"] >= 0], r["bd"][r["bi"] >= 0])
            nb = out["bird_counts"][j]
            assert out["bird_kps"][j][:nb].tobytes() == r["bk"].tobytes() and np.array_equal(out["bird_desc"][j][:nb], r["bdsc"])
            if r["m12"] is None:                      # first frame of an unchained call: nothing to match against
                assert out["bird_nmatches"][j] == 0 and (out["bird_matches12"][j] < 0).all()
            else:
                assert out["bird_nmatches"][j] == r["nbm"] and r["nbm"] > 20, (j, out["bird_nmatches"][j], r["nbm"])
                assert np.array_equal(out["bird_matches12"][j][:len(r["m12"])], r["m12"])
    # an unchained call forgets the carried frame
    out = step(seq["imgs"][0:4], seq["bird_imgs"][0:2], poses[0:2], chain=False)
    assert out["bird_nmatches"][0] == 0


def test_frame_step_one_frame_per_call_replayed_graph(pkg):
    """The drop-in pattern: ONE frame per orbb200_frame_step_host call from pageable memory, chained.  The first calls of a
    parameter set run eagerly, the next one is recorded into a CUDA graph (uploads, both front-ends, matching, downloads) and
    later ones replay it: every call, whichever way it ran, must equal the oracle -- keypoints, descriptors, mvuRight / mvDepth
    bits, SearchLocalPoints and BirdviewMatch results."""
    w, h, bw, bh, nfeat, bnf, nmap = 620, 188, 200, 200, 1000, 600, 1500
    seq = synth.northstar_sequence(7, 23, w=w, h=h, bird=(bw, bh), vehicle=(40, 60))
    orc = oracle.Extractor(nfeat, 1.2, 8, 20, 7)
    mp = synth.northstar_map(seq, lambda im: orc(im), nmap, 3)
    ctx = pkg.Context(nfeat, 1.2, 8, 20, 7, w, h, 2)
    M = pkg.LocalMap(ctx, mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"], mp["desc"])
    step = pkg.FrameStep(ctx, w, h, M, mb=0.537, mbf=386.1448, th=1.0, nnratio=0.8, bird_size=(bw, bh), bird_nfeatures=bnf,
                         bird_mask=seq["bird_mask"], bird_window=15, bird_nnratio=0.99)
    poses = [pkg.CameraPose.make(**p) for p in seq["poses"]]
    prev = None
    launches = []
    for i in range(7):
        before = ctx.launches
        out = step(seq["imgs"][2 * i:2 * i + 2], seq["bird_imgs"][i:i + 1], poses[i:i + 1], chain=i > 0)
        launches.append(ctx.launches - before)
        (r,), prev = _oracle_frame_step(seq, mp, [i], nfeat, bnf, w, h, bw, bh, 1.0, 0.8, 15, 0.99, prev)
        nl, nr = out["counts"][0], out["counts"][1]
        assert out["kps"][0][:nl].tobytes() == r["kl"].tobytes() and out["kps"][1][:nr].tobytes() == r["kr"].tobytes(), i
        assert np.array_equal(out["desc"][0][:nl], r["dl"]) and np.array_equal(out["desc"][1][:nr], r["dr"]), i
        assert np.array_equal(out["u_right"][0][:nl].view(np.uint32), r["ur"].view(np.uint32)), i
        assert np.array_equal(out["depth"][0][:nl].view(np.uint32), r["dep"].view(np.uint32)), i
        assert out["map_nmatches"][0] == r["nm"] and np.array_equal(out["map_best_idx"][0], r["bi"]), i
        assert np.array_equal(out["map_best_dist"][0][r["bi"] >= 0], r["bd"][r["bi"] >= 0]), i
        nb = out["bird_counts"][0]
        assert out["bird_kps"][0][:nb].tobytes() == r["bk"].tobytes() and np.array_equal(out["bird_desc"][0][:nb], r["bdsc"]), i
        if r["m12"] is None:
            assert out["bird_nmatches"][0] == 0
        else:
            assert out["bird_nmatches"][0] == r["nbm"] and np.array_equal(out["bird_matches12"][0][:len(r["m12"])], r["m12"]), i
    # eager, recorded and replayed calls enqueue the same kernels
    assert len(set(launches[1:])) == 1 and launches[1] > 30, launches


def test_frame_step_recording_failure_falls_back(pkg, monkeypatch):
    """When the recording of a frame step cannot be turned into a graph, the call that tried runs eagerly instead (nothing has
    executed at that point) and the parameter set is never recorded again: results stay equal to the oracle call after call."""
    monkeypatch.setenv("ORBB200_TEST_CAPTURE_FAIL", "1")
    w, h, bw, bh, nfeat, bnf, nmap = 620, 188, 200, 200, 1000, 600, 1500
    seq = synth.northstar_sequence(6, 31, w=w, h=h, bird=(bw, bh), vehicle=(40, 60))
    orc = oracle.Extractor(nfeat, 1.2, 8, 20, 7)
    mp = synth.northstar_map(seq, lambda im: orc(im), nmap, 3)
    ctx = pkg.Context(nfeat, 1.2, 8, 20, 7, w, h, 2)
    M = pkg.LocalMap(ctx, mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"], mp["desc"])
    step = pkg.FrameStep(ctx, w, h, M, mb=0.537, mbf=386.1448, th=1.0, nnratio=0.8, bird_size=(bw, bh), bird_nfeatures=bnf,
                         bird_mask=seq["bird_mask"], bird_window=15, bird_nnratio=0.99)
    poses = [pkg.CameraPose.make(**p) for p in seq["poses"]]
    prev = None
    for i in range(6):
        out = step(seq["imgs"][2 * i:2 * i + 2], seq["bird_imgs"][i:i + 1], poses[i:i + 1], chain=i > 0)
        (r,), prev = _oracle_frame_step(seq, mp, [i], nfeat, bnf, w, h, bw, bh, 1.0, 0.8, 15, 0.99, prev)
        nl, nb = out["counts"][0], out["bird_counts"][0]
        assert out["kps"][0][:nl].tobytes() == r["kl"].tobytes() and np.array_equal(out["desc"][0][:nl], r["dl"]), i
        assert out["map_nmatches"][0] == r["nm"] and np.array_equal(out["map_best_idx"][0], r["bi"]), i
        assert out["bird_kps"][0][:nb].tobytes() == r["bk"].tobytes() and np.array_equal(out["bird_desc"][0][:nb], r["bdsc"]), i
        if r["m12"] is not None:
            assert out["bird_nmatches"][0] == r["nbm"] and np.array_equal(out["bird_matches12"][0][:len(r["m12"])], r["m12"]), i


def test_frame_step_replay_survives_reallocation(pkg):
    """A recorded frame step holds device and pinned pointers.  Replacing the local map, changing the birdview mask and growing
    the staging blocks (a larger matcher call, a batch of two) between replayed calls must drop the recording, not replay it on
    stale pointers: every call still equals the oracle."""
    w, h, bw, bh, nfeat, bnf, nmap = 620, 188, 200, 200, 1000, 600, 1500
    seq = synth.northstar_sequence(12, 29, w=w, h=h, bird=(bw, bh), vehicle=(40, 60))
    orc = oracle.Extractor(nfeat, 1.2, 8, 20, 7)
    mp = synth.northstar_map(seq, lambda im: orc(im), nmap, 3)
    mp2 = synth.northstar_map(seq, lambda im: orc(im), nmap, 4)
    mask2 = seq["bird_mask"].copy()
    mask2[:40, :] = 0
    ctx = pkg.Context(nfeat, 1.2, 8, 20, 7, w, h, 4)
    poses = [pkg.CameraPose.make(**p) for p in seq["poses"]]

    def check(step, i, m, mask, prev):
        out = step(seq["imgs"][2 * i:2 * i + 2], seq["bird_imgs"][i:i + 1], poses[i:i + 1], chain=prev is not None)
        s2 = dict(seq, bird_mask=mask)
        (r,), prev = _oracle_frame_step(s2, m, [i], nfeat, bnf, w, h, bw, bh, 1.0, 0.8, 15, 0.99, prev)
        nl, nb = out["counts"][0], out["bird_counts"][0]
        assert out["kps"][0][:nl].tobytes() == r["kl"].tobytes() and np.array_equal(out["desc"][0][:nl], r["dl"]), i
        assert out["map_nmatches"][0] == r["nm"] and np.array_equal(out["map_best_idx"][0], r["bi"]), i
        assert out["bird_kps"][0][:nb].tobytes() == r["bk"].tobytes() and np.array_equal(out["bird_desc"][0][:nb], r["bdsc"]), i
        if r["m12"] is not None:
            assert out["bird_nmatches"][0] == r["nbm"] and np.array_equal(out["bird_matches12"][0][:len(r["m12"])], r["m12"]), i
        return prev

    def make(m, mask):
        M = pkg.LocalMap(ctx, m["pos"], m["normal"], m["max_distance"], m["min_distance"], m["desc"])
        return M, pkg.FrameStep(ctx, w, h, M, mb=0.537, mbf=386.1448, th=1.0, nnratio=0.8, bird_size=(bw, bh), bird_nfeatures=bnf,
                                bird_mask=mask, bird_window=15, bird_nnratio=0.99)
    M, step = make(mp, seq["bird_mask"])
    prev = None
    for i in range(5):                                # eager, eager, recorded, replayed, replayed
        prev = check(step, i, mp, seq["bird_mask"], prev)
    # a large matcher call grows the scratch blocks; a two-frame step uses other plans
    k, d = orc(seq["imgs"][0])
    F = pkg.Frame(ctx, k, d, 0.0, 0.0, 64.0 / w, 48.0 / h)
    nq = 60000
    rng = np.random.default_rng(3)
    idx = rng.integers(0, len(k), nq)
    pkg.ORBmatcher(ctx, 0.8).SearchByProjection(F, np.ones(nq, np.uint8), k["x"][idx], k["y"][idx], np.full(nq, -1, np.float32),
                                                k["octave"][idx].astype(np.int32), np.full(nq, 0.9, np.float32), d[idx].copy())
    for i in range(5, 8):
        prev = check(step, i, mp, seq["bird_mask"], prev)
    # new map (the old one freed: its addresses may be handed out again) and a new mask
    del step, M
    M, step = make(mp2, mask2)
    prev = None
    for i in range(8, 12):
        prev = check(step, i, mp2, mask2, prev)


def test_three_host_threads_three_contexts(pkg):
    """The reference calls the front-end from three threads at once (Tracking: extraction + projection searches; LocalMapping:
    SearchForTriangulation, src/LocalMapping.cc:278; LoopClosing: SearchByBoW, src/LoopClosing.cc:265), one ORBextractor /
    ORBmatcher each.  Three host threads with their own context on ONE GPU, running concurrently for a few hundred calls, must
    each reproduce the oracle exactly."""
    import threading
    imgs = [synth.synth_frame(240, 320, 900 + i) for i in range(4)]
    orc = oracle.Extractor(500, 1.2, 8, 20, 7)
    want_ex = [orc(im) for im in imgs]
    t = cases.triangulation_case(800, 800, 620, 188, 61, n_nodes=40)
    want_tri = oracle.search_for_triangulation(t["k1"], t["d1"], t["uR1"], t["has1"], t["k2"], t["d2"], t["uR2"], t["has2"], t["fv1"], t["fv2"],
                                               t["F12"], t["ex"], t["ey"], t["sf2"], t["sigma2"], False, True)
    b = cases.bow_case(800, 800, 620, 188, 81, n_nodes=40)
    O2 = oracle.Frame(b["k2"], b["d2"], b["grid"]["min_x"], b["grid"]["min_y"], b["grid"]["inv_w"], b["grid"]["inv_h"])
    want_bow = oracle.search_by_bow(b["d1"], b["k1"]["angle"], b["valid1"], O2, b["valid2"], b["fv1"], b["fv2"], 0.75, True, False)
    errors, start = [], threading.Barrier(3)
    reps = 60

    def tracking():
        try:
            ex = pkg.ORBextractor(500, 1.2, 8, 20, 7, max_size=(320, 240))
            start.wait()
            for r in range(reps):
                k, d = ex(imgs[r % 4])
                k0, d0 = want_ex[r % 4]
                assert k.tobytes() == k0.tobytes() and np.array_equal(d, d0), f"extract rep {r}"
        except Exception as e:            # noqa: BLE001
            errors.append(("tracking", repr(e)))

    def local_mapping():
        try:
            ctx = pkg.Context(2000, 1.2, 8, 20, 7, 64, 64)
            m = pkg.ORBmatcher(ctx, 0.6, True)
            start.wait()
            for r in range(reps):
                n, pairs = m.SearchForTriangulation(t["k1"], t["d1"], t["uR1"], t["has1"], t["k2"], t["d2"], t["uR2"], t["has2"], t["fv1"], t["fv2"],
                                                    t["F12"], t["ex"], t["ey"], t["sf2"], t["sigma2"], False)
                assert n == want_tri[0] and np.array_equal(pairs, want_tri[1]), f"triangulation rep {r}"
        except Exception as e:            # noqa: BLE001
            errors.append(("local_mapping", repr(e)))

    def loop_closing():
        try:
            ctx = pkg.Context(2000, 1.2, 8, 20, 7, 64, 64)
            F2 = pkg.Frame(ctx, b["k2"], b["d2"], b["grid"]["min_x"], b["grid"]["min_y"], b["grid"]["inv_w"], b["grid"]["inv_h"])
            m = pkg.ORBmatcher(ctx, 0.75, True)
            start.wait()
            for r in range(reps):
                nm, out = m.SearchByBoW(b["d1"], b["k1"]["angle"], b["valid1"], F2, b["fv1"], b["fv2"], b["valid2"], False)
                assert nm == want_bow[0] and np.array_equal(out, want_bow[1]), f"bow rep {r}"
        except Exception as e:            # noqa: BLE001
            errors.append(("loop_closing", repr(e)))

    threads = [threading.Thread(target=f) for f in (tracking, local_mapping, loop_closing)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert not errors, errors
    assert want_tri[0] > 0 and want_bow[0] > 0


def test_eight_extraction_threads_match_oracle(pkg):
    """The drop-in call from many threads (the reference extracts L and R on two threads, src/Frame.cc:124-127; tools/ubench/
    concurrent_calls.cu measures 1-8): eight host threads, one extractor (= one context) each on ONE GPU, every thread calls
    operator() on one image per call -- the one-graph staged path, eight graphs replayed concurrently -- and every result of every
    call must be the oracle's, byte for byte."""
    import threading
    shapes = [(240, 320, 500), (188, 620, 800)]
    imgs = {sh: [synth.synth_frame(sh[0], sh[1], 1300 + 10 * k + i) for i in range(3)] for k, sh in enumerate(shapes)}
    want = {sh: [oracle.Extractor(sh[2], 1.2, 8, 20, 7)(im) for im in imgs[sh]] for sh in shapes}
    T, reps = 8, 30
    errors, start = [], threading.Barrier(T)

    def worker(t):
        try:
            sh = shapes[t % 2]
            ex = pkg.ORBextractor(sh[2], 1.2, 8, 20, 7, max_size=(sh[1], sh[0]))
            start.wait()
            for r in range(reps):
                i = (r + t) % 3
                k, d = ex(imgs[sh][i])
                k0, d0 = want[sh][i]
                assert k.tobytes() == k0.tobytes() and np.array_equal(d, d0), f"thread {t} rep {r}"
        except Exception as e:            # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert not errors, errors
    assert all(len(k0) > 100 for sh in shapes for k0, _ in want[sh])


def test_two_contexts_different_feature_budgets(pkg):
    """the dynamic shared-memory opt-in is per (device, kernel); a second context with a smaller node table must not
    lower the first one's limit."""
    img = synth.synth_frame(376, 1241, 31)
    big = pkg.ORBextractor(8000, 1.2, 8, 20, 7, max_size=(1241, 376))
    k1, d1 = big(img)
    small = pkg.ORBextractor(500, 1.2, 8, 20, 7, max_size=(1241, 376))
    small(img)
    k2, d2 = big(img)
    assert k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2)
    k0, d0 = oracle.Extractor(8000, 1.2, 8, 20, 7)(img)
    assert k1.tobytes() == k0.tobytes() and np.array_equal(d1, d0)
