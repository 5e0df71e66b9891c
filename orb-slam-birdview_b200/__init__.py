"""orb-slam-birdview_b200 -- B200-native ORB front-end (extraction + Hamming matching).

Python host side over the C ABI of liborbb200.so (include/orbb200.h).  The classes mirror the reference's
C++ interface for the hot path -- ORB_SLAM2::ORBextractor (include/ORBextractor.h:44-111) and
ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:38-120) -- on numpy arrays: same names, argument meaning and
error behaviour.  There is no CPU path: importing works anywhere, but every compute call needs the CUDA
library and a device and raises OrbB200Error otherwise.
"""
import ctypes as C
import os

import numpy as np

__all__ = ["ORBextractor", "ORBmatcher", "ORBVocabulary", "Frame", "FrameStep", "LocalMap", "BirdviewORB", "Context", "OrbB200Error", "KP_DTYPE", "load_library", "build"]

_HERE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orb-slam-birdview_b200") \
    if os.path.basename(os.path.dirname(os.path.abspath(__file__))) != "orb-slam-birdview_b200" else os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liborbb200.so")

# == cv::KeyPoint, 28 bytes
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])

TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30           # src/ORBmatcher.cc:37-39
FRAME_GRID_ROWS, FRAME_GRID_COLS = 48, 64             # include/Frame.h:39-40


class OrbB200Error(RuntimeError):
    pass


def build(force=False, verbose=False):
    """Compile the CUDA extension in-tree (nvcc, sm_100a)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("orbb200_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)


_lib = None

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_SIGNATURES = {
    "orbb200_create": (_i, [C.POINTER(_vp), _i, _i, _f, _i, _i, _i, _i, _i, _i]),
    "orbb200_destroy": (None, [_vp]),
    "orbb200_last_error": (C.c_char_p, [_vp]),
    "orbb200_sync": (_i, [_vp]),
    "orbb200_stream": (_vp, [_vp]),
    "orbb200_get_levels": (_i, [_vp]),
    "orbb200_get_scale_table": (_i, [_vp, _i, _vp]),
    "orbb200_get_features_per_level": (_i, [_vp, _vp]),
    "orbb200_max_keypoints": (_i, [_vp]),
    "orbb200_extract": (_i, [_vp, _vp, _i, _i, _sz, _vp, _vp, _i, C.POINTER(_i)]),
    "orbb200_extract_batch": (_i, [_vp, C.POINTER(_vp), _i, _i, _i, _sz, _vp, _vp, _i, _vp]),
    "orbb200_extract_device": (_i, [_vp, _vp, _sz, _i, _i, _i, _sz]),
    "orbb200_results_device": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i)]),
    "orbb200_download_results": (_i, [_vp, _i, _vp, _vp, _i, _vp]),
    "orbb200_pyramid_level": (_i, [_vp, _i, _i, _i, _vp, _sz, C.POINTER(_i), C.POINTER(_i)]),
    "orbb200_pyramid_mirror": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "orbb200_set_pyramid_mirror": (_i, [_vp, _i]),
    "orbb200_level_candidates": (_i, [_vp, _i, _i, _vp, _i]),
    "orbb200_hamming_knn2": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "orbb200_distinctive_descriptors": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "orbb200_hamming_knn2_device": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "orbb200_measure_popc_peak": (C.c_double, [_vp]),
    "orbb200_frame_upload": (_i, [_vp, C.POINTER(_vp), _vp, _vp, _vp, _i, _f, _f, _f, _f]),
    "orbb200_frame_from_extract": (_i, [_vp, C.POINTER(_vp), _i, _f, _f, _f, _f]),
    "orbb200_frame_free": (None, [_vp]),
    "orbb200_frame_features_in_area": (_i, [_vp, _vp, _f, _f, _f, _i, _i, _vp, _i]),
    "orbb200_search_by_projection": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, C.POINTER(_i)]),
    "orbb200_search_by_projection_frame": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _i, _vp, C.POINTER(_i)]),
    "orbb200_birdview_match": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _f, _i, _vp, C.POINTER(_i)]),
    "orbb200_search_by_match_bird_kf": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _f, _f, _i, _vp, C.POINTER(_i)]),
    "orbb200_search_by_projection_bird": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, C.POINTER(_i)]),
    "orbb200_search_for_triangulation": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i,
                                              _vp, _f, _f, _vp, _vp, _i, _i, _vp, C.POINTER(_i)]),
    "orbb200_search_for_initialization": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _f, _i, _vp, C.POINTER(_i)]),
    "orbb200_search_window_best": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, C.POINTER(_i)]),
    "orbb200_search_by_bow": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _f, _i, _i, _vp, C.POINTER(_i)]),
    "orbb200_stereo_matches_device": (_i, [_vp, _i, _i, _i, _i, _f, _f]),
    "orbb200_stereo_results_device": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "orbb200_compute_stereo_matches": (_i, [_vp, _i, _i, _f, _f, _vp, _vp, _i, C.POINTER(_i)]),
    "orbb200_frame_from_extract_stereo": (_i, [_vp, C.POINTER(_vp), _i, _f, _f, _f, _f]),
    "orbb200_voc_create": (_i, [_vp, C.POINTER(_vp), _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "orbb200_voc_free": (None, [_vp]),
    "orbb200_bow_transform": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, C.POINTER(_i), _vp, _vp, _vp, C.POINTER(_i)]),
    "orbb200_bow_transform_extracted": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, C.POINTER(_i), _vp, _vp, _vp, C.POINTER(_i)]),
    "orbb200_bird_max_keypoints": (_i, [_vp, _i, _i, _i]),
    "orbb200_bird_detect": (_i, [_vp, _vp, _vp, _i, _i, _sz, _sz, _i, _vp, _i, C.POINTER(_i)]),
    "orbb200_corner_subpix": (_i, [_vp, _vp, _i, _i, _sz, _vp, _i, _i, _i, _i, C.c_double]),
    "orbb200_bird_compute": (_i, [_vp, _vp, _i, _i, _sz, _vp, _i, _vp, C.POINTER(_i)]),
    "orbb200_bird_extract": (_i, [_vp, _vp, _vp, _i, _i, _sz, _sz, _i, _vp, _vp, _i, C.POINTER(_i)]),
    "orbb200_bird_extract_batch": (_i, [_vp, _vp, _vp, _i, _i, _i, _sz, _sz, _i, _vp, _vp, _i, _vp]),
    "orbb200_map_upload": (_i, [_vp, C.POINTER(_vp), _i, _vp, _vp, _vp, _vp, _vp]),
    "orbb200_map_free": (None, [_vp]),
    "orbb200_is_in_frustum": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i)]),
    "orbb200_search_local_points": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i),
                                         _vp, _vp, _vp, C.POINTER(_i)]),
    "orbb200_stereo_step_device": (_i, [_vp, _vp, _sz, _i, _i, _i, _sz, _vp, _i, _f, _f, _f, _f, _f, _f, _vp, _vp, _vp]),
    "orbb200_stereo_step_host": (_i, [_vp, _vp, _i, _i, _i, _sz, _vp, _i, _f, _f, _f, _f, _f, _f, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "orbb200_step_enable_stereo": (_i, [_vp, _i, _f, _f]),
    "orbb200_stage_timing": (_i, [_vp, _i]),
    "orbb200_stage_times": (_i, [_vp, _vp, _vp, _i]),
    "orbb200_launch_count": (C.c_longlong, [_vp]),
    "orbb200_bird_set_mask": (_i, [_vp, _i, _i, _i, _i, _vp, _sz]),
    "orbb200_frame_step_device": (_i, [_vp, _vp, _vp, _vp]),
    "orbb200_frame_step_host": (_i, [_vp, _vp, _vp, _vp]),
    "orbb200_bird_results_device": (_i, [_vp, _i, _i, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i)]),
    "orbb200_device_status": (_i, [_vp, C.POINTER(_i)]),
}


class ProjQueries(C.Structure):
    """orbb200_proj_queries: device pointers of [n_frames][nq] query arrays"""
    _fields_ = [(n, C.c_void_p) for n in ("q_valid", "q_u", "q_v", "q_uR", "q_level", "q_viewcos", "q_desc", "q_obs_pos")]


class FrameStepParams(C.Structure):
    """orbb200_frame_step_params"""
    _fields_ = [("n_frames", C.c_int), ("w", C.c_int), ("h", C.c_int), ("stride", C.c_size_t), ("mb", C.c_float), ("mbf", C.c_float),
                ("min_x", C.c_float), ("min_y", C.c_float), ("inv_w", C.c_float), ("inv_h", C.c_float), ("map", C.c_void_p),
                ("viewing_cos_limit", C.c_float), ("th", C.c_float), ("nnratio", C.c_float),
                ("bird_w", C.c_int), ("bird_h", C.c_int), ("bird_stride", C.c_size_t), ("bird_nfeatures", C.c_int),
                ("bird_window", C.c_int), ("bird_nnratio", C.c_float), ("bird_check_ori", C.c_int), ("chain", C.c_int)]


class FrameStepInputs(C.Structure):
    """orbb200_frame_step_inputs"""
    _fields_ = [("imgs", C.c_void_p), ("bird_imgs", C.c_void_p), ("poses", C.c_void_p)]


class FrameStepOutputs(C.Structure):
    """orbb200_frame_step_outputs"""
    _fields_ = [(n, C.c_void_p) for n in ("kps", "desc", "counts", "u_right", "depth", "map_best_idx", "map_best_dist", "map_nmatches",
                                          "bird_kps", "bird_desc", "bird_counts", "bird_matches12", "bird_nmatches")] + \
               [("cap", C.c_int), ("bird_cap", C.c_int)]


class CameraPose(C.Structure):
    """orbb200_camera_pose: the Frame members Frame::isInFrustum reads (src/Frame.cc:436-492)"""
    _fields_ = [("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Ow", C.c_float * 3),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("mbf", C.c_float),
                ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float),
                ("log_scale_factor", C.c_float), ("n_levels", C.c_int32)]

    @classmethod
    def make(cls, Rcw, tcw, Ow, fx, fy, cx, cy, mbf, min_x, max_x, min_y, max_y, log_scale_factor, n_levels):
        p = cls()
        p.Rcw[:] = [float(v) for v in np.asarray(Rcw, np.float32).reshape(9)]
        p.tcw[:] = [float(v) for v in np.asarray(tcw, np.float32).reshape(3)]
        p.Ow[:] = [float(v) for v in np.asarray(Ow, np.float32).reshape(3)]
        p.fx, p.fy, p.cx, p.cy, p.mbf = fx, fy, cx, cy, mbf
        p.min_x, p.max_x, p.min_y, p.max_y = min_x, max_x, min_y, max_y
        p.log_scale_factor, p.n_levels = log_scale_factor, n_levels
        return p


def load_library():
    """dlopen liborbb200.so; fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OrbB200Error(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                           "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


class Context:
    """orbb200_ctx: one per host thread; owns a CUDA stream and the HBM-resident pools."""

    def __init__(self, nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, max_w, max_h, max_batch=1, device=0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.orbb200_create(C.byref(h), device, nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, max_w, max_h, max_batch)
        if rc != 0:
            raise OrbB200Error(f"orbb200_create failed ({rc}): {self._L.orbb200_last_error(None).decode()}")
        self._h = h
        self.nlevels, self.max_batch, self.device = nlevels, max_batch, device
        self.max_keypoints = self._L.orbb200_max_keypoints(h)

    def close(self):
        if getattr(self, "_h", None):
            self._L.orbb200_destroy(self._h)
            self._h = None

    __del__ = close

    def check(self, rc, what=""):
        if rc < 0:
            raise OrbB200Error(f"{what} failed ({rc}): {self._L.orbb200_last_error(self._h).decode()}")
        return rc

    def sync(self):
        self.check(self._L.orbb200_sync(self._h), "sync")

    @property
    def launches(self):
        return self._L.orbb200_launch_count(self._h)

    def scale_table(self, which):
        out = np.empty(self.nlevels, np.float32)
        self.check(self._L.orbb200_get_scale_table(self._h, which, _p(out)))
        return out

    def features_per_level(self):
        out = np.empty(self.nlevels, np.int32)
        self.check(self._L.orbb200_get_features_per_level(self._h, _p(out)))
        return out

    def popc_peak_gops(self):
        return self._L.orbb200_measure_popc_peak(self._h)


class ORBextractor:
    """Mirror of ORB_SLAM2::ORBextractor (reference include/ORBextractor.h:44-111).

    `extractor(image, mask=None)` is operator()(image, mask, keypoints, descriptors): returns
    (keypoints[KP_DTYPE], descriptors[N,32] uint8).  The mask is ignored, as in the reference
    (include/ORBextractor.h:58).  One instance per camera/thread, as in the reference."""

    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_size=(1920, 1080), max_batch=1, device=0):
        self.ctx = Context(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_size[0], max_size[1], max_batch, device)
        self._L = self.ctx._L
        self._scaleFactor = scaleFactor
        self.last_n = 0

    # -- getters (include/ORBextractor.h:63-82) --
    def GetLevels(self):
        return self.ctx.nlevels

    def GetScaleFactor(self):
        return self._scaleFactor

    def GetScaleFactors(self):
        return self.ctx.scale_table(0)

    def GetInverseScaleFactors(self):
        return self.ctx.scale_table(1)

    def GetScaleSigmaSquares(self):
        return self.ctx.scale_table(2)

    def GetInverseScaleSigmaSquares(self):
        return self.ctx.scale_table(3)

    def __call__(self, image, mask=None):
        if image is None or image.size == 0:
            return np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)     # empty image: silent return (:1046-1047)
        assert image.dtype == np.uint8 and image.ndim == 2, "image.type() == CV_8UC1"   # (:1050)
        k, d, n = self.extract_batch([image])
        return k[0, :n[0]].copy(), d[0, :n[0]].copy()

    def extract_batch(self, images):
        """n same-shape images -> (kps [n,cap], desc [n,cap,32], counts [n])"""
        n = len(images)
        h, w = images[0].shape
        imgs = []
        for im in images:
            assert im.dtype == np.uint8 and im.shape == (h, w)
            imgs.append(im if im.strides[1] == 1 else np.ascontiguousarray(im))
        stride = imgs[0].strides[0]
        imgs = [im if im.strides[0] == stride else np.ascontiguousarray(im) for im in imgs]
        stride = imgs[0].strides[0] if all(im.strides[0] == imgs[0].strides[0] for im in imgs) else None
        if stride is None:
            imgs = [np.ascontiguousarray(im) for im in imgs]
            stride = w
        cap = self.ctx.max_keypoints
        kps = np.empty((n, cap), KP_DTYPE)
        desc = np.empty((n, cap, 32), np.uint8)
        counts = np.empty(n, np.int32)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        self.ctx.check(self._L.orbb200_extract_batch(self.ctx._h, ptrs, n, w, h, stride, _p(kps), _p(desc), cap, _p(counts)), "extract")
        self.last_n = n
        return kps, desc, counts

    def ComputeStereoMatches(self, img_left, img_right, mb, mbf):
        """Frame::ComputeStereoMatches (src/Frame.cc:662-836) on two images of the last extract_batch call
        -> (n_matches, mvuRight[cap], mvDepth[cap])"""
        cap = self.ctx.max_keypoints
        ur, dp = np.empty(cap, np.float32), np.empty(cap, np.float32)
        nm = C.c_int()
        self.ctx.check(self._L.orbb200_compute_stereo_matches(self.ctx._h, img_left, img_right, mb, mbf, _p(ur), _p(dp), cap, C.byref(nm)),
                       "ComputeStereoMatches")
        return nm.value, ur, dp

    def image_pyramid(self, img_index=0, blurred=False):
        """mvImagePyramid (include/ORBextractor.h:85) of the last extraction, downloaded on demand."""
        out = []
        for lvl in range(self.ctx.nlevels):
            w, h = C.c_int(), C.c_int()
            self.ctx.check(self._L.orbb200_pyramid_level(self.ctx._h, img_index, lvl, int(blurred), None, 0, C.byref(w), C.byref(h)))
            a = np.empty((max(h.value, 0), max(w.value, 0)), np.uint8)
            if a.size:
                self.ctx.check(self._L.orbb200_pyramid_level(self.ctx._h, img_index, lvl, int(blurred), _p(a), a.strides[0], None, None))
            out.append(a)
        return out

    @property
    def mvImagePyramid(self):
        return self.image_pyramid(0, False)

    def level_candidates(self, img_index, level):
        n = self.ctx.check(self._L.orbb200_level_candidates(self.ctx._h, img_index, level, None, 0))
        out = np.empty((max(n, 1), 3), np.int32)
        self.ctx.check(self._L.orbb200_level_candidates(self.ctx._h, img_index, level, _p(out), n))
        return out[:n]


class Frame:
    """Device-resident flattened Frame: keypoints, descriptors, 64x48 lookup grid
    (Frame::AssignFeaturesToGrid, reference src/Frame.cc:378-412)."""

    def __init__(self, ctx, kps=None, desc=None, min_x=0.0, min_y=0.0, inv_w=1.0, inv_h=1.0, u_right=None, from_extract=None):
        self.ctx = ctx
        self._L = ctx._L
        h = C.c_void_p()
        if from_extract is not None:
            ctx.check(self._L.orbb200_frame_from_extract(ctx._h, C.byref(h), int(from_extract), min_x, min_y, inv_w, inv_h), "frame_from_extract")
            self.n = ctx.max_keypoints
        else:
            self.kps, self.desc = _c(kps, KP_DTYPE), _c(desc, np.uint8)
            self.u_right = _c(u_right, np.float32)
            self.n = len(self.kps)
            ctx.check(self._L.orbb200_frame_upload(ctx._h, C.byref(h), _p(self.kps), _p(self.desc), _p(self.u_right), self.n,
                                                   min_x, min_y, inv_w, inv_h), "frame_upload")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.orbb200_frame_free(self._h)
            self._h = None

    __del__ = close

    def GetFeaturesInArea(self, x, y, r, minLevel=-1, maxLevel=-1):
        out = np.empty(max(self.n, 1), np.int32)
        n = self.ctx.check(self._L.orbb200_frame_features_in_area(self.ctx._h, self._h, x, y, r, minLevel, maxLevel, _p(out), len(out)))
        return out[:n].copy()


class ORBVocabulary:
    """Flattened DBoW2 vocabulary (include/ORBVocabulary.h:30-31) with the transform used by Frame::ComputeBoW."""

    def __init__(self, ctx, child_ptr, child_idx, node_desc, word_id, weight, L):
        self.ctx, self._L = ctx, ctx._L
        a = [_c(child_ptr, np.int32), _c(child_idx, np.int32), _c(node_desc, np.uint8), _c(word_id, np.int32), _c(weight, np.float64)]
        h = C.c_void_p()
        ctx.check(self._L.orbb200_voc_create(ctx._h, C.byref(h), len(a[3]), *[_p(x) for x in a], int(L)), "voc_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.orbb200_voc_free(self._h)
            self._h = None

    __del__ = close

    def transform(self, desc=None, levelsup=4, img_index=None):
        """transform(features, BowVector, FeatureVector, levelsup)
        -> (word[n], node[n], (bow_word, bow_value), (fv_node, fv_ptr, fv_idx)); desc=None + img_index uses the descriptors
        of that image of the last extraction."""
        n = self.ctx.max_keypoints if desc is None else len(desc)
        word, node = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.int32)
        bw, bv = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.float64)
        fn, fp, fi = np.empty(max(n, 1), np.int32), np.empty(n + 1, np.int32), np.empty(max(n, 1), np.int32)
        nw, nf = C.c_int(), C.c_int()
        if desc is None:
            self.ctx.check(self._L.orbb200_bow_transform_extracted(self.ctx._h, self._h, int(img_index), int(levelsup), _p(word), _p(node), _p(bw),
                                                                   _p(bv), C.byref(nw), _p(fn), _p(fp), _p(fi), C.byref(nf)), "bow_transform")
        else:
            desc = _c(desc, np.uint8)
            self.ctx.check(self._L.orbb200_bow_transform(self.ctx._h, self._h, _p(desc), n, int(levelsup), _p(word), _p(node), _p(bw), _p(bv),
                                                         C.byref(nw), _p(fn), _p(fp), _p(fi), C.byref(nf)), "bow_transform")
        nfeat = int(fp[nf.value]) if nf.value > 0 else 0
        return word, node, (bw[:nw.value].copy(), bv[:nw.value].copy()), (fn[:nf.value].copy(), fp[:nf.value + 1].copy(), fi[:nfeat].copy())


class BirdviewORB:
    """The birdview front-end of the reference (src/Frame.cc:328-342): cv::ORB::create(nfeatures) detect(mask) +
    cv::cornerSubPix(5x5, 40 it, 1e-3) + compute, with cv::ORB's method names.  Results equal OpenCV 4.x's own C++ code
    (cv2 with setUseOptimized(False)) bit for bit, keypoint order included."""

    def __init__(self, ctx, nfeatures=2000):
        self.ctx, self._L, self.nfeatures = ctx, ctx._L, int(nfeatures)

    def _cap(self, w, h):
        cap = self._L.orbb200_bird_max_keypoints(self.ctx._h, w, h, self.nfeatures)
        if cap <= 0:
            raise OrbB200Error("bird_max_keypoints: " + self._L.orbb200_last_error(self.ctx._h).decode())
        return cap

    def detect(self, image, mask=None):
        img, msk = _c(image, np.uint8), _c(mask, np.uint8)
        h, w = img.shape
        cap = self._cap(w, h)
        kps = np.empty(cap, KP_DTYPE)
        n = C.c_int()
        self.ctx.check(self._L.orbb200_bird_detect(self.ctx._h, _p(img), _p(msk), w, h, img.strides[0], 0 if msk is None else msk.strides[0],
                                                   self.nfeatures, _p(kps), cap, C.byref(n)), "bird_detect")
        return kps[:n.value].copy()

    def cornerSubPix(self, image, pts, winSize=(5, 5), maxCount=40, epsilon=0.001):
        img = _c(image, np.uint8)
        pts = np.array(pts, np.float32, copy=True).reshape(-1, 2)
        self.ctx.check(self._L.orbb200_corner_subpix(self.ctx._h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(pts), len(pts),
                                                     int(winSize[0]), int(winSize[1]), int(maxCount), float(epsilon)), "cornerSubPix")
        return pts

    def compute(self, image, kps):
        img = _c(image, np.uint8)
        kps = np.array(kps, KP_DTYPE, copy=True)
        desc = np.empty((max(len(kps), 1), 32), np.uint8)
        n = C.c_int()
        self.ctx.check(self._L.orbb200_bird_compute(self.ctx._h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), len(kps), _p(desc),
                                                    C.byref(n)), "bird_compute")
        return kps[:n.value].copy(), desc[:n.value].copy()

    def __call__(self, image, mask=None):
        """detect + cornerSubPix + compute on the device -> (mvKeysBird, mDescriptorsBird)"""
        k, d = self.extract_batch([image], None if mask is None else [mask])
        return k[0], d[0]

    def extract_batch(self, images, masks=None):
        imgs = [_c(i, np.uint8) for i in images]
        msks = None if masks is None else [_c(m, np.uint8) for m in masks]
        n = len(imgs)
        h, w = imgs[0].shape
        if any(i.shape != (h, w) or i.strides[0] != imgs[0].strides[0] for i in imgs):
            raise ValueError("images of one batch must share shape and stride")
        cap = self._cap(w, h)
        kps = np.empty((n, cap), KP_DTYPE)
        desc = np.empty((n, cap, 32), np.uint8)
        cnt = np.empty(n, np.int32)
        ip = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
        mp = None if msks is None else (C.c_void_p * n)(*[m.ctypes.data for m in msks])
        self.ctx.check(self._L.orbb200_bird_extract_batch(self.ctx._h, ip, mp, n, w, h, imgs[0].strides[0], 0 if msks is None else msks[0].strides[0],
                                                          self.nfeatures, _p(kps), _p(desc), cap, _p(cnt)), "bird_extract")
        return [kps[i, :cnt[i]].copy() for i in range(n)], [desc[i, :cnt[i]].copy() for i in range(n)]


class LocalMap:
    """Device-resident snapshot of Tracking::mvpLocalMapPoints: world position, normal, mfMaxDistance / mfMinDistance
    and descriptor per map point.  isInFrustum mirrors Frame::isInFrustum over the whole map (src/Frame.cc:436-492)."""

    def __init__(self, ctx, pos, normal, max_distance, min_distance, desc):
        self.ctx, self._L = ctx, ctx._L
        a = [_c(pos, np.float32), _c(normal, np.float32), _c(max_distance, np.float32), _c(min_distance, np.float32), _c(desc, np.uint8)]
        self.n = len(a[2])
        h = C.c_void_p()
        ctx.check(self._L.orbb200_map_upload(ctx._h, C.byref(h), self.n, *[_p(x) for x in a]), "map_upload")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.orbb200_map_free(self._h)
            self._h = None

    __del__ = close

    def _outs(self):
        n = max(self.n, 1)
        return (np.empty(n, np.uint8), np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32),
                np.empty(n, np.int32), np.empty(n, np.float32))

    def isInFrustum(self, pose, viewingCosLimit=0.5, candidate=None):
        """-> (nToMatch, mbTrackInView, mTrackProjX, mTrackProjY, mTrackProjXR, mnTrackScaleLevel, mTrackViewCos)"""
        iv, u, v, uR, lvl, vc = self._outs()
        cand = _c(candidate, np.uint8)
        k = C.c_int()
        self.ctx.check(self._L.orbb200_is_in_frustum(self.ctx._h, self._h, C.addressof(pose), viewingCosLimit, _p(cand), _p(iv), _p(u), _p(v),
                                                     _p(uR), _p(lvl), _p(vc), C.byref(k)), "isInFrustum")
        n = self.n
        return k.value, iv[:n], u[:n], v[:n], uR[:n], lvl[:n], vc[:n]


class FrameStep:
    """The batched north-star frame (orbb200_frame_step_host): per frame L+R ORBextractor, ComputeStereoMatches, birdview
    cv::ORB + cornerSubPix + compute, SearchLocalPoints against a device-resident local map and SearchByMatchBird against the
    previous frame -- what Frame::Frame (src/Frame.cc:84-142, 263-375) and Tracking::Track (src/Tracking.cc:1241, 1610-1660)
    run on the data-parallel side, for n frames per call on numpy (host) buffers."""

    def __init__(self, ctx, w, h, local_map=None, mb=0.0, mbf=0.0, grid=None, th=1.0, nnratio=0.8, viewing_cos_limit=0.5,
                 bird_size=None, bird_nfeatures=2000, bird_mask=None, bird_window=15, bird_nnratio=0.99, bird_check_ori=True):
        self.ctx, self._L = ctx, ctx._L
        self.w, self.h, self.map = w, h, local_map
        P = FrameStepParams()
        P.w, P.h, P.stride, P.mb, P.mbf = w, h, w, mb, mbf
        g = grid or (0.0, 0.0, FRAME_GRID_COLS / w, FRAME_GRID_ROWS / h)
        P.min_x, P.min_y, P.inv_w, P.inv_h = g
        P.map = local_map._h if local_map is not None else None
        P.viewing_cos_limit, P.th, P.nnratio = viewing_cos_limit, th, nnratio
        if bird_size:
            P.bird_w, P.bird_h = bird_size
            P.bird_stride, P.bird_nfeatures = bird_size[0], bird_nfeatures
            P.bird_window, P.bird_nnratio, P.bird_check_ori = bird_window, bird_nnratio, int(bird_check_ori)
            self.bird_cap = self._L.orbb200_bird_max_keypoints(ctx._h, P.bird_w, P.bird_h, bird_nfeatures)
            m = _c(bird_mask, np.uint8)
            ctx.check(self._L.orbb200_bird_set_mask(ctx._h, P.bird_w, P.bird_h, bird_nfeatures, max(ctx.max_batch // 2, 1), _p(m),
                                                    m.strides[0] if m is not None else 0), "bird_set_mask")
        else:
            self.bird_cap = 0
        self.P = P
        self.cap = ctx.max_keypoints

    def __call__(self, imgs, bird_imgs=None, poses=None, chain=False):
        """imgs [2n][h][w] (left, right interleaved), bird_imgs [n][bh][bw], poses: list of n CameraPose.
        -> dict of host arrays (kps, desc, counts, u_right, depth, map_best_idx, map_best_dist, map_nmatches, bird_kps, bird_desc,
        bird_counts, bird_matches12, bird_nmatches)"""
        imgs = np.ascontiguousarray(imgs, np.uint8)
        n = imgs.shape[0] // 2
        P = self.P
        P.n_frames, P.chain = n, int(chain)
        I, O = FrameStepInputs(), FrameStepOutputs()
        I.imgs = imgs.ctypes.data
        out = dict(kps=np.zeros((2 * n, self.cap), KP_DTYPE), desc=np.zeros((2 * n, self.cap, 32), np.uint8), counts=np.zeros(2 * n, np.int32))
        if P.mb > 0:
            out.update(u_right=np.zeros((n, self.cap), np.float32), depth=np.zeros((n, self.cap), np.float32))
        keep = [imgs]
        if self.map is not None:
            arr = (CameraPose * n)(*poses)
            keep.append(arr)
            I.poses = C.addressof(arr)
            out.update(map_best_idx=np.zeros((n, self.map.n), np.int32), map_best_dist=np.zeros((n, self.map.n), np.int32),
                       map_nmatches=np.zeros(n, np.int32))
        if P.bird_w > 0:
            b = np.ascontiguousarray(bird_imgs, np.uint8)
            keep.append(b)
            I.bird_imgs = b.ctypes.data
            out.update(bird_kps=np.zeros((n, self.bird_cap), KP_DTYPE), bird_desc=np.zeros((n, self.bird_cap, 32), np.uint8),
                       bird_counts=np.zeros(n, np.int32), bird_matches12=np.full((n, self.bird_cap), -1, np.int32),
                       bird_nmatches=np.zeros(n, np.int32))
        for k, v in out.items():
            setattr(O, k, v.ctypes.data)
        O.cap, O.bird_cap = self.cap, self.bird_cap
        self.ctx.check(self._L.orbb200_frame_step_host(self.ctx._h, C.byref(P), C.byref(I), C.byref(O)), "frame_step_host")
        self.ctx.sync()
        st = C.c_int()
        self.ctx.check(self._L.orbb200_device_status(self.ctx._h, C.byref(st)), "device_status")
        return out


class ORBmatcher:
    """Mirror of ORB_SLAM2::ORBmatcher (reference include/ORBmatcher.h:38-120) on flattened inputs.

    The reference walks MapPoint*/KeyFrame* graphs; here the caller passes the same data as arrays (the
    C++ shim in cpp/ does that flattening for the real classes).  Method names and arguments follow the
    reference; results are returned instead of written into Frame members."""

    TH_LOW, TH_HIGH, HISTO_LENGTH = TH_LOW, TH_HIGH, HISTO_LENGTH

    def __init__(self, ctx, nnratio=0.6, checkOri=True):
        self.ctx = ctx
        self._L = ctx._L
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)

    @staticmethod
    def DescriptorDistance(a, b):
        raise OrbB200Error("single-pair DescriptorDistance stays on the host in the reference shim; use knn2/search calls")

    def hamming_knn2(self, q, m):
        q, m = _c(q, np.uint8), _c(m, np.uint8)
        nq, nm = len(q), len(m)
        bi, bd, sd = (np.empty(nq, np.int32) for _ in range(3))
        self.ctx.check(self._L.orbb200_hamming_knn2(self.ctx._h, _p(q), nq, _p(m), nm, _p(bi), _p(bd), _p(sd)), "hamming_knn2")
        return bi, bd, sd

    def ComputeDistinctiveDescriptors(self, desc, group_ptr):
        """Selection step of MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:272-301) and
        MapPointBird::ComputeDistinctiveDescriptors (src/MapPointBird.cc:117-146) for many landmarks at once:
        descriptors of landmark g's observations are desc[group_ptr[g]:group_ptr[g+1]]
        -> (best index inside each group or -1, its median distance)"""
        desc, group_ptr = _c(desc, np.uint8), _c(group_ptr, np.int32)
        ng = len(group_ptr) - 1
        bi, bm = np.empty(max(ng, 0), np.int32), np.empty(max(ng, 0), np.int32)
        self.ctx.check(self._L.orbb200_distinctive_descriptors(self.ctx._h, _p(desc), _p(group_ptr), ng, _p(bi), _p(bm)),
                       "distinctive_descriptors")
        return bi, bm

    def SearchByProjection(self, F, q_valid, q_u, q_v, q_uR, q_level, q_viewcos, q_desc, q_obs_pos=None, kp_blocked=None, th=1.0):
        """SearchByProjection(Frame&, vector<MapPoint*>&, th) (src/ORBmatcher.cc:45-129)
        -> (nmatches, best_idx[nq], best_dist[nq], query_of_kp[n])"""
        nq = len(q_u)
        a = [_c(q_valid, np.uint8), _c(q_u, np.float32), _c(q_v, np.float32), _c(q_uR, np.float32), _c(q_level, np.int32),
             _c(q_viewcos, np.float32), _c(q_desc, np.uint8), _c(q_obs_pos, np.uint8), _c(kp_blocked, np.uint8)]
        bi, bd = np.empty(nq, np.int32), np.empty(nq, np.int32)
        qk = np.full(max(F.n, 1), -1, np.int32)
        nm = C.c_int()
        self.ctx.check(self._L.orbb200_search_by_projection(self.ctx._h, F._h, nq, *[_p(x) for x in a], th, self.mfNNratio,
                                                            _p(bi), _p(bd), _p(qk), C.byref(nm)), "SearchByProjection")
        return nm.value, bi, bd, qk[:F.n]

    def SearchLocalPoints(self, F, local_map, pose, candidate=None, obs_pos=None, kp_blocked=None, th=1.0, viewingCosLimit=0.5):
        """Tracking::SearchLocalPoints (src/Tracking.cc:1634-1660): isInFrustum over the device-resident local map, then
        SearchByProjection(F, local map, th).  -> (nToMatch, in_view, (u, v, uR, level, viewcos), nmatches, best_idx, best_dist,
        query_of_kp)"""
        iv, u, v, uR, lvl, vc = local_map._outs()
        n = local_map.n
        a = [_c(candidate, np.uint8), _c(obs_pos, np.uint8), _c(kp_blocked, np.uint8)]
        bi, bd = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.int32)
        qk = np.full(max(F.n, 1), -1, np.int32)
        k, nm = C.c_int(), C.c_int()
        self.ctx.check(self._L.orbb200_search_local_points(self.ctx._h, F._h, local_map._h, C.addressof(pose), viewingCosLimit, *[_p(x) for x in a],
                                                           th, self.mfNNratio, _p(iv), _p(u), _p(v), _p(uR), _p(lvl), _p(vc), C.byref(k),
                                                           _p(bi), _p(bd), _p(qk), C.byref(nm)), "SearchLocalPoints")
        return k.value, iv[:n], (u[:n], v[:n], uR[:n], lvl[:n], vc[:n]), nm.value, bi[:n], bd[:n], qk[:F.n]

    def SearchByProjectionFrame(self, Cur, q_valid, q_u, q_v, q_invz, q_octave, q_angle, q_desc, q_obs_pos=None, kp_blocked=None,
                                th=15.0, mbf=0.0, mode=0):
        """SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) (src/ORBmatcher.cc:1328-1470)"""
        nq = len(q_u)
        a = [_c(q_valid, np.uint8), _c(q_u, np.float32), _c(q_v, np.float32), _c(q_invz, np.float32), _c(q_octave, np.int32),
             _c(q_angle, np.float32), _c(q_desc, np.uint8), _c(q_obs_pos, np.uint8), _c(kp_blocked, np.uint8)]
        qk = np.full(max(Cur.n, 1), -1, np.int32)
        nm = C.c_int()
        self.ctx.check(self._L.orbb200_search_by_projection_frame(self.ctx._h, Cur._h, nq, *[_p(x) for x in a], th, mbf, mode,
                                                                  int(self.mbCheckOrientation), _p(qk), C.byref(nm)), "SearchByProjection")
        return nm.value, qk[:Cur.n]

    def BirdviewMatch(self, kps1, desc1, F2, windowSize, vPrevMatched=None):
        """BirdviewMatch (src/ORBmatcher.cc:1667-1786 with vPrevMatched, :1788-1899 without)
        -> (nmatches, vnMatches12, vPrevMatched')"""
        kps1, desc1 = _c(kps1, KP_DTYPE), _c(desc1, np.uint8)
        prev = None if vPrevMatched is None else np.array(vPrevMatched, np.float32, copy=True).reshape(-1, 2)
        m12 = np.full(max(len(kps1), 1), -1, np.int32)
        nm = C.c_int()
        self.ctx.check(self._L.orbb200_birdview_match(self.ctx._h, _p(kps1), _p(desc1), len(kps1), F2._h, _p(prev), int(windowSize),
                                                      self.mfNNratio, int(self.mbCheckOrientation), _p(m12), C.byref(nm)), "BirdviewMatch")
        return nm.value, m12[:len(kps1)], prev

    def SearchByMatchBird(self, kf_kps, has_mp, mp_desc, F, r):
        """SearchByMatchBird(KeyFrame*, Frame&, out, r) (src/ORBmatcher.cc:2000-2114)"""
        kf_kps, has_mp, mp_desc = _c(kf_kps, KP_DTYPE), _c(has_mp, np.uint8), _c(mp_desc, np.uint8)
        out = np.full(max(F.n, 1), -1, np.int32)
        nm = C.c_int()
        self.ctx.check(self._L.orbb200_search_by_match_bird_kf(self.ctx._h, _p(kf_kps), _p(has_mp), _p(mp_desc), len(kf_kps), F._h, r,
                                                               self.mfNNratio, int(self.mbCheckOrientation), _p(out), C.byref(nm)), "SearchByMatchBird")
        return nm.value, out[:F.n]

    def SearchByProjectionBird(self, F, q_valid, q_x, q_y, q_desc, q_obs_pos=None, kp_blocked=None, r=4.0):
        """SearchByProjectionBird(Frame&, vector<MapPointBird*>&, r) (src/ORBmatcher.cc:1923-1998)"""
        nq = len(q_x)
        a = [_c(q_valid, np.uint8), _c(q_x, np.float32), _c(q_y, np.float32), _c(q_desc, np.uint8), _c(q_obs_pos, np.uint8),
             _c(kp_blocked, np.uint8)]
        out = np.full(max(F.n, 1), -1, np.int32)
        nm = C.c_int()
        self.ctx.check(self._L.orbb200_search_by_projection_bird(self.ctx._h, F._h, nq, *[_p(x) for x in a], r, self.mfNNratio,
                                                                 _p(out), C.byref(nm)), "SearchByProjectionBird")
        return nm.value, out[:F.n]

    def SearchForTriangulation(self, kps1, desc1, uR1, has_mp1, kps2, desc2, uR2, has_mp2, fv1, fv2, F12, ex, ey,
                               scale_factors2, level_sigma2_2, bOnlyStereo=False):
        """SearchForTriangulation(KF1, KF2, F12, pairs, bOnlyStereo) (src/ORBmatcher.cc:657-823)"""
        kps1, kps2 = _c(kps1, KP_DTYPE), _c(kps2, KP_DTYPE)
        desc1, desc2 = _c(desc1, np.uint8), _c(desc2, np.uint8)
        uR1, uR2 = _c(uR1, np.float32), _c(uR2, np.float32)
        has_mp1, has_mp2 = _c(has_mp1, np.uint8), _c(has_mp2, np.uint8)
        f1 = [_c(x, np.int32) for x in fv1]
        f2 = [_c(x, np.int32) for x in fv2]
        F12 = _c(F12, np.float32).reshape(9)
        sf2, ls2 = _c(scale_factors2, np.float32), _c(level_sigma2_2, np.float32)
        pairs = np.empty((max(len(kps1), 1), 2), np.int32)
        npairs = C.c_int()
        self.ctx.check(self._L.orbb200_search_for_triangulation(
            self.ctx._h, _p(kps1), _p(desc1), _p(uR1), _p(has_mp1), len(kps1), _p(kps2), _p(desc2), _p(uR2), _p(has_mp2), len(kps2),
            _p(f1[0]), _p(f1[1]), _p(f1[2]), len(f1[0]), _p(f2[0]), _p(f2[1]), _p(f2[2]), len(f2[0]),
            _p(F12), ex, ey, _p(sf2), _p(ls2), int(bOnlyStereo), int(self.mbCheckOrientation), _p(pairs), C.byref(npairs)),
            "SearchForTriangulation")
        return npairs.value, pairs[:npairs.value].copy()

    WB_BLOCK, WB_URCHECK, WB_CHI2, WB_ORI = 1, 2, 4, 8

    def SearchForInitialization(self, kps1, desc1, F2, vbPrevMatched, windowSize=100):
        """SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:405-520)"""
        kps1, desc1 = _c(kps1, KP_DTYPE), _c(desc1, np.uint8)
        prev = np.array(vbPrevMatched, np.float32, copy=True).reshape(-1, 2)
        m12 = np.full(max(len(kps1), 1), -1, np.int32)
        nm = C.c_int()
        self.ctx.check(self._L.orbb200_search_for_initialization(self.ctx._h, _p(kps1), _p(desc1), len(kps1), F2._h, _p(prev), int(windowSize),
                                                                 self.mfNNratio, int(self.mbCheckOrientation), _p(m12), C.byref(nm)),
                       "SearchForInitialization")
        return nm.value, m12[:len(kps1)], prev

    def search_window_best(self, F, q_valid, q_x, q_y, q_r, q_minL, q_maxL, q_desc, q_aux=None, q_angle=None, q_obs_pos=None,
                           kp_blocked=None, inv_level_sigma2=None, acc_th=50, flags=0):
        """Generic best-in-window search behind SearchByProjection(Frame,KF,set) / (KF,Scw), Fuse, SearchBySim3
        (include/orbb200.h: orbb200_search_window_best)."""
        nq = len(q_x)
        a = [_c(q_valid, np.uint8), _c(q_x, np.float32), _c(q_y, np.float32), _c(q_r, np.float32), _c(q_minL, np.int32), _c(q_maxL, np.int32),
             _c(q_desc, np.uint8), _c(q_aux, np.float32), _c(q_angle, np.float32), _c(q_obs_pos, np.uint8), _c(kp_blocked, np.uint8),
             _c(inv_level_sigma2, np.float32)]
        bi, bd = np.empty(nq, np.int32), np.empty(nq, np.int32)
        qk = np.full(max(F.n, 1), -1, np.int32)
        nm = C.c_int()
        self.ctx.check(self._L.orbb200_search_window_best(self.ctx._h, F._h, nq, *[_p(x) for x in a], int(acc_th), int(flags),
                                                          _p(bi), _p(bd), _p(qk), C.byref(nm)), "search_window_best")
        return nm.value, bi, bd, qk[:F.n]

    def SearchByBoW(self, desc1, angle1, valid1, F2, fv1, fv2, valid2=None, kf_kf=False):
        """SearchByBoW(KeyFrame*, Frame&, ..) (src/ORBmatcher.cc:159-288) / (KeyFrame*, KeyFrame*, ..) (:522-655)"""
        desc1, angle1, valid1, valid2 = _c(desc1, np.uint8), _c(angle1, np.float32), _c(valid1, np.uint8), _c(valid2, np.uint8)
        f1 = [_c(x, np.int32) for x in fv1]
        f2 = [_c(x, np.int32) for x in fv2]
        n1 = len(desc1)
        out = np.full(max(n1 if kf_kf else F2.n, 1), -1, np.int32)
        nm = C.c_int()
        self.ctx.check(self._L.orbb200_search_by_bow(self.ctx._h, _p(desc1), _p(angle1), _p(valid1), n1, F2._h, _p(valid2),
                                                     _p(f1[0]), _p(f1[1]), _p(f1[2]), len(f1[0]), _p(f2[0]), _p(f2[1]), _p(f2[2]), len(f2[0]),
                                                     self.mfNNratio, int(self.mbCheckOrientation), int(kf_kf), _p(out), C.byref(nm)), "SearchByBoW")
        return nm.value, out[:(n1 if kf_kf else F2.n)]
