"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liborb_oracle.so, the dependency-free CPU restatement of the reference's ORB
front-end (oracle/orb_oracle.h).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(native=False, force=False, out_dir=None):
    """Compile the oracle.  native=True -> -march=native copy (bench CPU baseline, built on the box)."""
    name = "liborb_oracle_native.so" if native else "liborb_oracle.so"
    out = os.path.join(out_dir or _HERE, name)
    srcs = [os.path.join(_HERE, f) for f in ("orb_oracle.cpp", "match_oracle.cpp", "bird_oracle.cpp", "orb_oracle.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "orbb200_pattern.inc"))
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    cmd = [os.environ.get("CXX", "g++"), "-O3", "-march=native" if native else "-march=x86-64-v3", "-std=c++17",
           "-ffp-contract=off", "-fPIC", "-shared", "-o", out,
           os.path.join(_HERE, "orb_oracle.cpp"), os.path.join(_HERE, "match_oracle.cpp"), os.path.join(_HERE, "bird_oracle.cpp")]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return out


_libs = {}


def lib(native=False):
    if native in _libs:
        return _libs[native]
    path = build(native=native)
    L = C.CDLL(path)
    u8p, i32p, f32p, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_void_p
    sz = C.c_size_t
    sig = {
        "oracle_resize_u8": (None, [vp, C.c_int, C.c_int, sz, vp, C.c_int, C.c_int, sz]),
        "oracle_gauss7_u8": (None, [vp, C.c_int, C.c_int, sz, vp, sz]),
        "oracle_fast9": (C.c_int, [vp, C.c_int, C.c_int, sz, C.c_int, C.c_int, vp, C.c_int]),
        "oracle_fast_score_map": (None, [vp, C.c_int, C.c_int, sz, vp]),
        "oracle_fast_atan2": (C.c_float, [C.c_float, C.c_float]),
        "oracle_cv_round": (C.c_int, [C.c_float]),
        "oracle_extractor_create": (vp, [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]),
        "oracle_extractor_destroy": (None, [vp]),
        "oracle_extract": (C.c_int, [vp, vp, C.c_int, C.c_int, sz, vp, vp, C.c_int]),
        "oracle_extractor_stage_ms": (None, [vp, vp, C.c_int]),
        "oracle_extractor_features_per_level": (C.c_int, [vp, vp]),
        "oracle_extractor_scale_factors": (C.c_int, [vp, vp]),
        "oracle_extractor_umax": (C.c_int, [vp, vp]),
        "oracle_extractor_level_size": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "oracle_extractor_level_image": (C.c_int, [vp, C.c_int, C.c_int, vp, sz]),
        "oracle_extractor_level_candidates": (C.c_int, [vp, C.c_int, vp, C.c_int]),
        "oracle_extractor_level_keypoints": (C.c_int, [vp, C.c_int, vp, C.c_int]),
        "oracle_compute_stereo_matches": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_float, C.c_float, vp, vp]),
        "oracle_distribute_octree": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]),
        "oracle_descriptor_distance": (C.c_int, [vp, vp]),
        "oracle_hamming_knn2": (None, [vp, C.c_int, vp, C.c_int, vp, vp, vp]),
        "oracle_distinctive_descriptors": (None, [vp, vp, C.c_int, vp, vp]),
        "oracle_frame_create": (vp, [vp, vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, vp]),
        "oracle_frame_destroy": (None, [vp]),
        "oracle_frame_features_in_area": (C.c_int, [vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, vp, C.c_int]),
        "oracle_search_by_projection": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                                  C.c_float, C.c_float, vp, vp, vp]),
        "oracle_search_by_projection_frame": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                                        C.c_float, C.c_float, C.c_int, C.c_int, vp]),
        "oracle_birdview_match": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, C.c_float, C.c_int, vp]),
        "oracle_search_by_match_bird_kf": (C.c_int, [vp, vp, vp, C.c_int, vp, C.c_float, C.c_float, C.c_int, vp]),
        "oracle_search_by_projection_bird": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_float, C.c_float, vp]),
        "oracle_search_for_initialization": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, C.c_float, C.c_int, vp]),
        "oracle_search_window_best": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]),
        "oracle_search_by_bow": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.c_float, C.c_int, C.c_int, vp]),
        "oracle_voc_create": (vp, [C.c_int, vp, vp, vp, vp, vp, C.c_int]),
        "oracle_voc_destroy": (None, [vp]),
        "oracle_resize_linear_exact_u8": (None, [vp, C.c_int, C.c_int, sz, vp, C.c_int, C.c_int, sz]),
        "oracle_sep_gauss7_f32_u8": (None, [vp, C.c_int, C.c_int, sz, vp, sz]),
        "oracle_bird_detect": (C.c_int, [vp, vp, C.c_int, C.c_int, sz, sz, C.c_int, vp, C.c_int]),
        "oracle_get_rect_sub_pix_8u32f": (None, [vp, C.c_int, C.c_int, sz, C.c_float, C.c_float, C.c_int, C.c_int, vp]),
        "oracle_corner_subpix": (None, [vp, C.c_int, C.c_int, sz, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
        "oracle_bird_compute": (C.c_int, [vp, C.c_int, C.c_int, sz, vp, C.c_int, vp]),
        "oracle_bird_extract": (C.c_int, [vp, vp, C.c_int, C.c_int, sz, sz, C.c_int, vp, vp, C.c_int]),
        "oracle_is_in_frustum": (C.c_int, [C.c_int, vp, vp, vp, vp, vp, vp, C.c_float, vp, vp, vp, vp, vp, vp]),
        "oracle_voc_transform": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "oracle_search_for_triangulation": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int,
                                                      vp, vp, vp, C.c_int, vp, vp, vp, C.c_int,
                                                      vp, C.c_float, C.c_float, vp, vp, C.c_int, C.c_int, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _libs[native] = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


# ---- primitives --------------------------------------------------------------------------------
def resize_u8(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().oracle_resize_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dst.strides[0])
    return dst


def gauss7_u8(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().oracle_gauss7_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def fast9(img, threshold, nms=True):
    """-> int32 [n,3] (x, y, response) in cv::FAST output order"""
    assert img.dtype == np.uint8 and img.strides[1] == 1
    cap = max(16, img.shape[0] * img.shape[1])
    out = np.empty((cap, 3), np.int32)
    n = lib().oracle_fast9(C.c_void_p(img.ctypes.data), img.shape[1], img.shape[0], img.strides[0], threshold, int(nms), _p(out), cap)
    return out[:n].copy()


def fast_score_map(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(img.shape, np.int32)
    lib().oracle_fast_score_map(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out))
    return out


def fast_atan2(y, x):
    return lib().oracle_fast_atan2(float(y), float(x))


def cv_round(v):
    return lib().oracle_cv_round(float(v))


# ---- extractor ---------------------------------------------------------------------------------
class Extractor:
    """Mirror of ORB_SLAM2::ORBextractor (reference include/ORBextractor.h:44-111)."""

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7, native=False):
        self._L = lib(native)
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self._h = self._L.oracle_extractor_create(nfeatures, scale, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.oracle_extractor_destroy(self._h)
            self._h = None

    def __call__(self, img):
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        cap = self.nfeatures * 2 + 4096
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = self._L.oracle_extract(self._h, C.c_void_p(img.ctypes.data), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc), cap)
        if n < 0:
            raise RuntimeError("oracle_extract: capacity too small")
        return kps[:n].copy(), desc[:n].copy()

    def features_per_level(self):
        out = np.empty(self.nlevels, np.int32)
        self._L.oracle_extractor_features_per_level(self._h, _p(out))
        return out

    def stage_ms(self, reset=False):
        """accumulated wall ms per stage: dict(pyramid, fast, octree, orientation, blur, describe)"""
        out = np.zeros(6, np.float64)
        self._L.oracle_extractor_stage_ms(self._h, _p(out), int(reset))
        return dict(zip(("pyramid", "fast", "octree", "orientation", "blur", "describe"), out.tolist()))

    def scale_factors(self):
        out = np.empty(self.nlevels, np.float32)
        self._L.oracle_extractor_scale_factors(self._h, _p(out))
        return out

    def umax(self):
        out = np.empty(16, np.int32)
        self._L.oracle_extractor_umax(self._h, _p(out))
        return out

    def level_image(self, level, blurred=False):
        w, h = C.c_int(), C.c_int()
        self._L.oracle_extractor_level_size(self._h, level, C.byref(w), C.byref(h))
        out = np.empty((h.value, w.value), np.uint8)
        if out.size:
            self._L.oracle_extractor_level_image(self._h, level, int(blurred), _p(out), out.strides[0])
        return out

    def level_candidates(self, level):
        n = self._L.oracle_extractor_level_candidates(self._h, level, None, 0)
        out = np.empty((max(n, 1), 3), np.int32)
        self._L.oracle_extractor_level_candidates(self._h, level, _p(out), n)
        return out[:n]

    def level_keypoints(self, level):
        n = self._L.oracle_extractor_level_keypoints(self._h, level, None, 0)
        out = np.empty(max(n, 1), KP_DTYPE)
        self._L.oracle_extractor_level_keypoints(self._h, level, _p(out), n)
        return out[:n]


def distribute_octree(xyr, minX, maxX, minY, maxY, N):
    xyr = _c(xyr, np.int32).reshape(-1, 3)
    out = np.empty((max(len(xyr), 1), 3), np.int32)
    n = lib().oracle_distribute_octree(_p(xyr), len(xyr), minX, maxX, minY, maxY, N, _p(out), len(out))
    return out[:n].copy()


# ---- matcher -----------------------------------------------------------------------------------
def descriptor_distance(a, b):
    a = _c(a, np.uint8)
    b = _c(b, np.uint8)
    return lib().oracle_descriptor_distance(_p(a), _p(b))


def hamming_knn2(q, m, native=False):
    q = _c(q, np.uint8)
    m = _c(m, np.uint8)
    nq, nm = len(q), len(m)
    bi, bd, sd = (np.empty(nq, np.int32) for _ in range(3))
    lib(native).oracle_hamming_knn2(_p(q), nq, _p(m), nm, _p(bi), _p(bd), _p(sd))
    return bi, bd, sd


def distinctive_descriptors(desc, group_ptr, native=False):
    """MapPoint[Bird]::ComputeDistinctiveDescriptors selection (src/MapPoint.cc:272-301, src/MapPointBird.cc:117-146)
    -> (best index inside each group or -1, its median distance)"""
    desc, group_ptr = _c(desc, np.uint8), _c(group_ptr, np.int32)
    ng = len(group_ptr) - 1
    bi, bm = np.empty(ng, np.int32), np.empty(ng, np.int32)
    lib(native).oracle_distinctive_descriptors(_p(desc), _p(group_ptr), ng, _p(bi), _p(bm))
    return bi, bm


class Frame:
    """Flattened Frame: keypoints (undistorted), descriptors, 64x48 grid (reference src/Frame.cc:378-412)."""

    def __init__(self, kps, desc, min_x, min_y, inv_w, inv_h, u_right=None, native=False):
        self._L = lib(native)
        self.kps = _c(kps, KP_DTYPE)
        self.desc = _c(desc, np.uint8)
        self.n = len(self.kps)
        self.u_right = _c(u_right, np.float32)
        self._h = self._L.oracle_frame_create(_p(self.kps), _p(self.desc), self.n, min_x, min_y, inv_w, inv_h, _p(self.u_right))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.oracle_frame_destroy(self._h)
            self._h = None

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.empty(max(self.n, 1), np.int32)
        n = self._L.oracle_frame_features_in_area(self._h, x, y, r, min_level, max_level, _p(out), len(out))
        return out[:n].copy()


def search_by_projection(F, scale_factors, q_valid, q_u, q_v, q_uR, q_level, q_viewcos, q_desc, q_obs_pos=None,
                         kp_blocked=None, th=1.0, nnratio=0.8):
    nq = len(q_u)
    sf = _c(scale_factors, np.float32)
    a = [_c(q_valid, np.uint8), _c(q_u, np.float32), _c(q_v, np.float32), _c(q_uR, np.float32), _c(q_level, np.int32),
         _c(q_viewcos, np.float32), _c(q_desc, np.uint8), _c(q_obs_pos, np.uint8), _c(kp_blocked, np.uint8)]
    bi, bd = np.empty(nq, np.int32), np.empty(nq, np.int32)
    qk = np.empty(F.n, np.int32)
    n = F._L.oracle_search_by_projection(F._h, _p(sf), nq, *[_p(x) for x in a], th, nnratio, _p(bi), _p(bd), _p(qk))
    return n, bi, bd, qk


def search_by_projection_frame(Cur, scale_factors, q_valid, q_u, q_v, q_invz, q_octave, q_angle, q_desc, q_obs_pos=None,
                               kp_blocked=None, th=15.0, mbf=0.0, mode=0, check_ori=True):
    nq = len(q_u)
    sf = _c(scale_factors, np.float32)
    a = [_c(q_valid, np.uint8), _c(q_u, np.float32), _c(q_v, np.float32), _c(q_invz, np.float32), _c(q_octave, np.int32),
         _c(q_angle, np.float32), _c(q_desc, np.uint8), _c(q_obs_pos, np.uint8), _c(kp_blocked, np.uint8)]
    qk = np.empty(Cur.n, np.int32)
    n = Cur._L.oracle_search_by_projection_frame(Cur._h, _p(sf), nq, *[_p(x) for x in a], th, mbf, mode, int(check_ori), _p(qk))
    return n, qk


def birdview_match(kps1, desc1, F2, prev_xy=None, window=15, nnratio=0.99, check_ori=True):
    kps1 = _c(kps1, KP_DTYPE)
    desc1 = _c(desc1, np.uint8)
    prev = None if prev_xy is None else np.array(prev_xy, np.float32, copy=True).reshape(-1, 2)
    m12 = np.empty(len(kps1), np.int32)
    n = F2._L.oracle_birdview_match(_p(kps1), _p(desc1), len(kps1), F2._h, _p(prev), int(window), nnratio, int(check_ori), _p(m12))
    return n, m12, prev


def search_by_match_bird_kf(kf_kps, has_mp, mp_desc, F, r=15.0, nnratio=0.99, check_ori=True):
    kf_kps = _c(kf_kps, KP_DTYPE)
    has_mp = _c(has_mp, np.uint8)
    mp_desc = _c(mp_desc, np.uint8)
    out = np.empty(F.n, np.int32)
    n = F._L.oracle_search_by_match_bird_kf(_p(kf_kps), _p(has_mp), _p(mp_desc), len(kf_kps), F._h, r, nnratio, int(check_ori), _p(out))
    return n, out


def search_by_projection_bird(F, q_valid, q_x, q_y, q_desc, q_obs_pos=None, kp_blocked=None, r=4.0, nnratio=0.99):
    nq = len(q_x)
    a = [_c(q_valid, np.uint8), _c(q_x, np.float32), _c(q_y, np.float32), _c(q_desc, np.uint8), _c(q_obs_pos, np.uint8),
         _c(kp_blocked, np.uint8)]
    out = np.empty(F.n, np.int32)
    n = F._L.oracle_search_by_projection_bird(F._h, nq, *[_p(x) for x in a], r, nnratio, _p(out))
    return n, out


def search_for_triangulation(kps1, desc1, uR1, has_mp1, kps2, desc2, uR2, has_mp2, fv1, fv2, F12, ex, ey,
                             scale_factors2, level_sigma2_2, only_stereo=False, check_ori=False):
    """fv = (node ids ascending int32[nn], ptr int32[nn+1], idx int32[...])"""
    kps1, kps2 = _c(kps1, KP_DTYPE), _c(kps2, KP_DTYPE)
    desc1, desc2 = _c(desc1, np.uint8), _c(desc2, np.uint8)
    uR1, uR2 = _c(uR1, np.float32), _c(uR2, np.float32)
    has_mp1, has_mp2 = _c(has_mp1, np.uint8), _c(has_mp2, np.uint8)
    f1 = [_c(x, np.int32) for x in fv1]
    f2 = [_c(x, np.int32) for x in fv2]
    F12 = _c(F12, np.float32).reshape(9)
    sf2, ls2 = _c(scale_factors2, np.float32), _c(level_sigma2_2, np.float32)
    pairs = np.empty((max(len(kps1), 1), 2), np.int32)
    n = lib().oracle_search_for_triangulation(
        _p(kps1), _p(desc1), _p(uR1), _p(has_mp1), len(kps1), _p(kps2), _p(desc2), _p(uR2), _p(has_mp2), len(kps2),
        _p(f1[0]), _p(f1[1]), _p(f1[2]), len(f1[0]), _p(f2[0]), _p(f2[1]), _p(f2[2]), len(f2[0]),
        _p(F12), ex, ey, _p(sf2), _p(ls2), int(only_stereo), int(check_ori), _p(pairs))
    return n, pairs[:n].copy()


WB_BLOCK, WB_URCHECK, WB_CHI2, WB_ORI = 1, 2, 4, 8


def search_for_initialization(kps1, desc1, F2, prev_xy, window=100, nnratio=0.9, check_ori=True):
    kps1 = _c(kps1, KP_DTYPE)
    desc1 = _c(desc1, np.uint8)
    prev = np.array(prev_xy, np.float32, copy=True).reshape(-1, 2)
    m12 = np.empty(len(kps1), np.int32)
    n = F2._L.oracle_search_for_initialization(_p(kps1), _p(desc1), len(kps1), F2._h, _p(prev), int(window), nnratio, int(check_ori), _p(m12))
    return n, m12, prev


def search_window_best(F, q_valid, q_x, q_y, q_r, q_minL, q_maxL, q_desc, q_aux=None, q_angle=None, q_obs_pos=None, kp_blocked=None,
                       inv_level_sigma2=None, acc_th=50, flags=0):
    nq = len(q_x)
    a = [_c(q_valid, np.uint8), _c(q_x, np.float32), _c(q_y, np.float32), _c(q_r, np.float32), _c(q_minL, np.int32), _c(q_maxL, np.int32),
         _c(q_desc, np.uint8), _c(q_aux, np.float32), _c(q_angle, np.float32), _c(q_obs_pos, np.uint8), _c(kp_blocked, np.uint8),
         _c(inv_level_sigma2, np.float32)]
    bi, bd = np.empty(nq, np.int32), np.empty(nq, np.int32)
    qk = np.empty(F.n, np.int32)
    n = F._L.oracle_search_window_best(F._h, nq, *[_p(x) for x in a], int(acc_th), int(flags), _p(bi), _p(bd), _p(qk))
    return n, bi, bd, qk


def search_by_bow(desc1, angle1, valid1, F2, valid2, fv1, fv2, nnratio=0.7, check_ori=True, kf_kf=False):
    desc1, angle1, valid1 = _c(desc1, np.uint8), _c(angle1, np.float32), _c(valid1, np.uint8)
    valid2 = _c(valid2, np.uint8)
    f1 = [_c(x, np.int32) for x in fv1]
    f2 = [_c(x, np.int32) for x in fv2]
    n1 = len(desc1)
    out = np.empty(n1 if kf_kf else F2.n, np.int32)
    n = F2._L.oracle_search_by_bow(_p(desc1), _p(angle1), _p(valid1), n1, F2._h, _p(valid2), _p(f1[0]), _p(f1[1]), _p(f1[2]), len(f1[0]),
                                   _p(f2[0]), _p(f2[1]), _p(f2[2]), len(f2[0]), nnratio, int(check_ori), int(kf_kf), _p(out))
    return n, out


def compute_stereo_matches(ex_left, ex_right, kps_l, desc_l, kps_r, desc_r, mb, mbf):
    """Frame::ComputeStereoMatches; ex_left/ex_right are Extractor objects whose last call was on the left/right image"""
    kps_l, kps_r = _c(kps_l, KP_DTYPE), _c(kps_r, KP_DTYPE)
    desc_l, desc_r = _c(desc_l, np.uint8), _c(desc_r, np.uint8)
    ur = np.empty(len(kps_l), np.float32)
    dp = np.empty(len(kps_l), np.float32)
    n = ex_left._L.oracle_compute_stereo_matches(ex_left._h, ex_right._h, _p(kps_l), _p(desc_l), len(kps_l), _p(kps_r), _p(desc_r), len(kps_r),
                                                 mb, mbf, _p(ur), _p(dp))
    return n, ur, dp


class Vocabulary:
    """Flattened DBoW2 vocabulary tree: child_ptr/child_idx (CSR), node descriptors, leaf word ids (-1 inner), weights"""

    def __init__(self, child_ptr, child_idx, node_desc, word_id, weight, L):
        self._L = lib()
        self.child_ptr, self.child_idx = _c(child_ptr, np.int32), _c(child_idx, np.int32)
        self.node_desc, self.word_id, self.weight, self.L = _c(node_desc, np.uint8), _c(word_id, np.int32), _c(weight, np.float64), L
        self._h = self._L.oracle_voc_create(len(self.word_id), _p(self.child_ptr), _p(self.child_idx), _p(self.node_desc), _p(self.word_id),
                                            _p(self.weight), L)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.oracle_voc_destroy(self._h)
            self._h = None

    def transform(self, desc, levelsup=4):
        """-> (word[n], node[n], (bow_word, bow_value), (fv_node, fv_ptr, fv_idx))"""
        desc = _c(desc, np.uint8)
        n = len(desc)
        word, node = np.empty(n, np.int32), np.empty(n, np.int32)
        bw, bv = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.float64)
        fn, fp, fi = np.empty(max(n, 1), np.int32), np.empty(n + 1, np.int32), np.empty(max(n, 1), np.int32)
        nw, nf = C.c_int32(), C.c_int32()
        self._L.oracle_voc_transform(self._h, _p(desc), n, levelsup, _p(word), _p(node), _p(bw), _p(bv), C.byref(nw), _p(fn), _p(fp), _p(fi), C.byref(nf))
        return word, node, (bw[:nw.value].copy(), bv[:nw.value].copy()), (fn[:nf.value].copy(), fp[:nf.value + 1].copy(), fi[:fp[nf.value]].copy())


class CameraPose(C.Structure):
    """Frame pose + intrinsics read by Frame::isInFrustum (mirror of oracle_camera_pose / orbb200_camera_pose)."""
    _fields_ = [("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Ow", C.c_float * 3),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("mbf", C.c_float),
                ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float),
                ("log_scale_factor", C.c_float), ("n_levels", C.c_int32)]


def camera_pose(Rcw, tcw, Ow, fx, fy, cx, cy, mbf, min_x, max_x, min_y, max_y, log_scale_factor, n_levels, cls=CameraPose):
    p = cls()
    p.Rcw[:] = [float(v) for v in np.asarray(Rcw, np.float32).reshape(9)]
    p.tcw[:] = [float(v) for v in np.asarray(tcw, np.float32).reshape(3)]
    p.Ow[:] = [float(v) for v in np.asarray(Ow, np.float32).reshape(3)]
    p.fx, p.fy, p.cx, p.cy, p.mbf = fx, fy, cx, cy, mbf
    p.min_x, p.max_x, p.min_y, p.max_y = min_x, max_x, min_y, max_y
    p.log_scale_factor, p.n_levels = log_scale_factor, n_levels
    return p


def is_in_frustum(pos, normal, max_distance, min_distance, pose, viewing_cos_limit=0.5, candidate=None):
    """Frame::isInFrustum over an array of map points -> (nToMatch, in_view, u, v, uR, level, viewcos)."""
    L = lib()
    pos, normal = _c(pos, np.float32), _c(normal, np.float32)
    n = len(pos)
    mx, mn = _c(max_distance, np.float32), _c(min_distance, np.float32)
    cand = _c(candidate, np.uint8)
    iv = np.empty(n, np.uint8)
    u, v, uR, vc = (np.empty(n, np.float32) for _ in range(4))
    lvl = np.empty(n, np.int32)
    k = L.oracle_is_in_frustum(n, _p(pos), _p(normal), _p(mx), _p(mn), _p(cand), C.addressof(pose), viewing_cos_limit,
                               _p(iv), _p(u), _p(v), _p(uR), _p(lvl), _p(vc))
    return k, iv, u, v, uR, lvl, vc


# ---- birdview front-end: cv::ORB + cornerSubPix (reference src/Frame.cc:328-342) ---------------------------
def resize_linear_exact_u8(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().oracle_resize_linear_exact_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dst.strides[0])
    return dst


def sep_gauss7_f32_u8(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().oracle_sep_gauss7_f32_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def bird_detect(img, mask=None, nfeatures=2000):
    img = np.ascontiguousarray(img, np.uint8)
    mask = _c(mask, np.uint8)
    h, w = img.shape
    cap = 4 * nfeatures + 64
    out = np.empty(cap, KP_DTYPE)
    n = lib().oracle_bird_detect(_p(img), _p(mask), w, h, img.strides[0], 0 if mask is None else mask.strides[0], nfeatures, _p(out), cap)
    return out[:n].copy()


def get_rect_sub_pix(img, cx, cy, win_w, win_h):
    img = np.ascontiguousarray(img, np.uint8)
    dst = np.empty((win_h, win_w), np.float32)
    lib().oracle_get_rect_sub_pix_8u32f(_p(img), img.shape[1], img.shape[0], img.strides[0], cx, cy, win_w, win_h, _p(dst))
    return dst


def corner_subpix(img, pts, win=(5, 5), max_iter=40, eps=0.001):
    img = np.ascontiguousarray(img, np.uint8)
    pts = np.array(pts, np.float32, copy=True).reshape(-1, 2)
    lib().oracle_corner_subpix(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(pts), len(pts), win[0], win[1], max_iter, eps)
    return pts


def bird_compute(img, kps):
    img = np.ascontiguousarray(img, np.uint8)
    kps = np.array(kps, KP_DTYPE, copy=True)
    desc = np.empty((max(len(kps), 1), 32), np.uint8)
    n = lib().oracle_bird_compute(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), len(kps), _p(desc))
    return kps[:n].copy(), desc[:n].copy()


def bird_extract(img, mask=None, nfeatures=2000):
    img = np.ascontiguousarray(img, np.uint8)
    mask = _c(mask, np.uint8)
    h, w = img.shape
    cap = 4 * nfeatures + 64
    kps, desc = np.empty(cap, KP_DTYPE), np.empty((cap, 32), np.uint8)
    n = lib().oracle_bird_extract(_p(img), _p(mask), w, h, img.strides[0], 0 if mask is None else mask.strides[0], nfeatures, _p(kps), _p(desc), cap)
    return kps[:n].copy(), desc[:n].copy()


# ---- oracle/_ref: the REFERENCE's own ORBextractor.cc, compiled unmodified (oracle/ref_standin/Makefile) -----------
REF_DIR = os.path.join(_HERE, "_ref")
REF_SRC = os.environ.get("ORB_REFERENCE_ROOT", "/root/reference")
_REF_VARIANTS = {"glibc": "libref_orbextractor.so", "bump": "libref_orbextractor_bump.so", "nofma": "libref_orbextractor_nofma.so"}


def build_ref(force=False):
    """Build oracle/_ref/*.so from the reference sources where they lie.  Needs /root/reference (this container);
    on the GPU box the prebuilt files travel with the snapshot.  Returns True when the libraries exist."""
    have_src = os.path.exists(os.path.join(REF_SRC, "src", "ORBextractor.cc"))
    if have_src:
        cmd = ["make", "-s", "-C", os.path.join(_HERE, "ref_standin"), f"REF={REF_SRC}"]
        if force:
            cmd.insert(1, "-B")
        subprocess.run(cmd, check=True)
    return ref_available()


def ref_available(variant="glibc"):
    return os.path.exists(os.path.join(REF_DIR, _REF_VARIANTS[variant]))


_ref_libs = {}


def ref_lib(variant="glibc"):
    if variant in _ref_libs:
        return _ref_libs[variant]
    path = os.path.join(REF_DIR, _REF_VARIANTS[variant])
    if not os.path.exists(path):
        build_ref()
    L = C.CDLL(path)
    vp, sz = C.c_void_p, C.c_size_t
    sig = {
        "ref_is_bump_alloc": (C.c_int, []),
        "ref_extractor_create": (vp, [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]),
        "ref_extractor_destroy": (None, [vp]),
        "ref_extract": (C.c_int, [vp, vp, C.c_int, C.c_int, sz, vp, vp, C.c_int]),
        "ref_extractor_levels": (C.c_int, [vp]),
        "ref_extractor_scale_table": (C.c_int, [vp, C.c_int, vp]),
        "ref_extractor_features_per_level": (C.c_int, [vp, vp]),
        "ref_extractor_umax": (C.c_int, [vp, vp]),
        "ref_extractor_level_size": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "ref_extractor_level_image": (C.c_int, [vp, C.c_int, C.c_int, vp, sz]),
        "ref_distribute_octree": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _ref_libs[variant] = L
    return L


class RefExtractor:
    """ORB_SLAM2::ORBextractor of the reference itself (src/ORBextractor.cc compiled unmodified).
    variant: "glibc" = the real binary's allocator (heap-address tie-break of ORBextractor.cc:684 as glibc gives it,
    FMA contraction on like a -march=native build); "bump" = monotone operator new, no contraction (must equal
    oracle.Extractor byte for byte); "nofma" = glibc allocator, no contraction."""

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7, variant="glibc"):
        self._L = ref_lib(variant)
        self.nfeatures, self.nlevels, self.variant = nfeatures, nlevels, variant
        self._h = self._L.ref_extractor_create(nfeatures, scale, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_extractor_destroy(self._h)
            self._h = None

    def __call__(self, img):
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        cap = self.nfeatures * 2 + 4096
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = self._L.ref_extract(self._h, C.c_void_p(img.ctypes.data), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc), cap)
        if n < 0:
            raise RuntimeError("ref_extract: capacity too small")
        return kps[:n].copy(), desc[:n].copy()

    def features_per_level(self):
        out = np.empty(self.nlevels, np.int32)
        self._L.ref_extractor_features_per_level(self._h, _p(out))
        return out

    def scale_table(self, which=0):
        out = np.empty(self.nlevels, np.float32)
        self._L.ref_extractor_scale_table(self._h, which, _p(out))
        return out

    def scale_factors(self):
        return self.scale_table(0)

    def umax(self):
        out = np.empty(16, np.int32)
        self._L.ref_extractor_umax(self._h, _p(out))
        return out

    def level_image(self, level, border=0):
        w, h = C.c_int(), C.c_int()
        self._L.ref_extractor_level_size(self._h, level, C.byref(w), C.byref(h))
        out = np.empty((h.value + 2 * border, w.value + 2 * border), np.uint8)
        if out.size:
            self._L.ref_extractor_level_image(self._h, level, border, _p(out), out.strides[0])
        return out

    def distribute_octree(self, xyr, minX, maxX, minY, maxY, N):
        xyr = _c(xyr, np.int32).reshape(-1, 3)
        out = np.empty((max(len(xyr), 1), 3), np.int32)
        n = self._L.ref_distribute_octree(self._h, _p(xyr), len(xyr), minX, maxX, minY, maxY, N, _p(out), len(out))
        return out[:n].copy()
