#!/usr/bin/env python3
"""bench.py -- ORB front-end throughput on B200 (BASELINE.json metric: extract+match frames/s).

Headline workload (config C2, BASELINE.json configs[1]): KITTI-shape stereo frames, 1241x376 left+right, 2000 features
per image, 8 levels, scale 1.2, FAST 20/7; per frame: ORBextractor on both images, lookup grid on the left frame,
Tracking::SearchLocalPoints against a device-resident local map of 3000 points (Frame::isInFrustum per point, then
ORBmatcher::SearchByProjection, th=1, ratio 0.8).  One "step" = one pass of that hot path over a batch of
--frames-per-step synthetic stereo frames (a sequence: the camera moves between frames).

Also on the same JSON line, under "configs" (VERDICT r1):
  C3_full  the north-star frame: the C2 work + Frame::ComputeStereoMatches + the birdview front-end (cv::ORB(2000)
           detect(mask) + cornerSubPix + compute on a 400x400 image) + SearchByMatchBird against the previous frame;
  C1, C5   extraction only, 752x480/1000 and 1920x1080/4000 (C5 also as a sharded 256-frame sequence with a digest that
           must not depend on the number of GPUs);
  C4       brute-force Hamming 2k x {2k, 20k, 200k} with the fraction of the measured POPC issue peak;
  per_frame  the drop-in path: one frame per call through the host API.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU)
  python bench.py --impl reference --steps K --warmup W     the reference on the host cores: its own ORBextractor.cc
                                                            compiled unmodified (oracle/_ref) + the oracle port of the matcher

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events); `e2e` = the same step through the
host-buffer C-ABI call, H2D/D2H inside the timed region; `roofline` = the dominant kernel against the measured HBM peak;
`cpu_baseline` = the reference on the box's host cores (bounded sample).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, NFEAT, NLEVELS, SCALE, INI_TH, MIN_TH = 1241, 376, 2000, 8, 1.2, 20, 7
NMAP, TH, RATIO, COS_LIMIT = 3000, 1.0, 0.8, 0.5
BW, BH, BNFEAT, BIRD_WINDOW, BIRD_RATIO = 400, 400, 2000, 15, 0.99      # src/Frame.cc:329, src/Tracking.cc:326,1241
GRID = (0.0, 0.0, 64.0 / W, 48.0 / H)        # mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv (k1 == 0)
METRIC = "ORB extract+match frames/sec"
STAGES = ["import", "pyramid", "fast", "blur", "octree", "describe", "grid", "match", "stereo", "bird_pyramid", "bird_detect", "bird_subpix",
          "bird_describe", "frustum", "-", "-"]
NSTAGES = 16
MB, MBF = 0.537, 386.1448                      # KITTI00-02.yaml: Camera.bf / Camera.fx, Camera.bf


def _synth():
    from importlib import import_module
    return import_module("orb_slam_birdview_b200.synth")


def config_dict(B, P):
    """`config` of the JSON line: identical in both arms (the driver compares them)."""
    return {"workload": "C2: KITTI-shape stereo 1241x376 L+R, 2000 feats/img, 8 levels, scale 1.2, FAST 20/7, extract + "
                        "SearchLocalPoints (isInFrustum + SearchByProjection, th=1, ratio 0.8) vs 3000 local map points",
            "frames_per_step_per_gpu": B, "queries_per_frame": NMAP, "sharding": "frames over ranks (no collective on the path)",
            "l2": f"{P} rotating input batches of {2 * B} images + {2 * B}-image pyramid/blur pools: working set "
                  f"{working_set_mb(B, P):.0f} MB per GPU > 126 MB L2"}


def level_pixels(w, h, last=False):
    s = np.float32(1.0)
    tot, p0, pl = 0, w * h, 0
    for _ in range(NLEVELS):
        inv = np.float32(1.0) / s
        lw, lh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))
        tot += lw * lh
        pl = lw * lh
        s = np.float32(np.float64(s) * np.float64(np.float32(SCALE)))
    if last:
        return pl
    return tot, p0


def working_set_mb(B, P):
    sum_p, p0 = level_pixels(W, H)
    return (P * 2 * B * p0 + 2 * 2 * B * sum_p * 1.05) / 1e6


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def host_cpu_model():
    """CPU model of the box the CPU arm runs on (core count and model go into the line)."""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    import platform
    return platform.processor() or platform.machine()


def make_pool(B, seed, extract):
    """One input batch: a B-frame sequence (front L/R + birdview + poses) and its local map."""
    synth = _synth()
    seq = synth.northstar_sequence(B, seed, w=W, h=H, bird=(BW, BH))
    seq["map"] = synth.northstar_map(seq, extract, NMAP, seed + 7)
    return seq


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own ORBextractor.cc (oracle/_ref) + the oracle port -- baseline infrastructure, never the product
# ---------------------------------------------------------------------------------------------------------
class CpuArm:
    """Per-frame work of the two measured workloads on host threads.  Extractors are per thread and live as long as the
    arm (the reference keeps one ORBextractor per camera for the whole run, src/Tracking.cc:121-127); the thread pool is
    persistent across steps."""

    def __init__(self, threads, use_ref=True):
        import oracle
        self.oracle, self.threads = oracle, threads
        try:
            oracle.lib(native=True)          # -O3 -march=native, the reference's own flags (CMakeLists.txt:10-11)
            self.native = True
        except Exception:
            self.native = False
        self.use_ref = bool(use_ref and oracle.ref_available("glibc"))
        self.tls = threading.local()
        self.pool = None
        if threads > 1:
            from concurrent.futures import ThreadPoolExecutor
            self.pool = ThreadPoolExecutor(threads)
        self.sf = oracle.Extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, native=self.native).scale_factors()

    def close(self):
        if self.pool:
            self.pool.shutdown()
            self.pool = None

    def _port(self, k):
        """k-th oracle-port extractor of this thread (0 left, 1 right)"""
        if not hasattr(self.tls, "port"):
            self.tls.port = {}
        if k not in self.tls.port:
            self.tls.port[k] = self.oracle.Extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, native=self.native)
        return self.tls.port[k]

    def _ref(self):
        if not hasattr(self.tls, "ref"):
            self.tls.ref = self.oracle.RefExtractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, variant="glibc")
        return self.tls.ref

    def frame_c2(self, seq, i):
        """C2: extract L+R (the reference runs them on two threads; here a worker does both, all cores being busy with other
        frames), grid on the left frame, isInFrustum + SearchByProjection against the local map"""
        o = self.oracle
        ex = self._ref() if self.use_ref else self._port(0)
        kl, dl = ex(seq["imgs"][2 * i])
        kr, dr = ex(seq["imgs"][2 * i + 1])
        F = o.Frame(kl, dl, *[np.float32(g) for g in GRID], native=self.native)
        mp = seq["map"]
        k0, iv, u, v, uR, lvl, vc = o.is_in_frustum(mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"], seq["cpu_poses"][i], COS_LIMIT, None)
        nm, bi, bd, qk = o.search_by_projection(F, self.sf, iv, u, v, uR, lvl, vc, mp["desc"], None, None, TH, RATIO)
        return nm

    def frame_full(self, seq, i):
        """the north-star frame: C2 + ComputeStereoMatches + birdview front-end + BirdviewMatch(previous, current)"""
        o = self.oracle
        exl, exr = self._port(0), self._port(1)      # ComputeStereoMatches reads both extractors' pyramids (src/Frame.cc:669-776)
        kl, dl = exl(seq["imgs"][2 * i])
        kr, dr = exr(seq["imgs"][2 * i + 1])
        _, ur, _ = o.compute_stereo_matches(exl, exr, kl, dl, kr, dr, MB, MBF)
        F = o.Frame(kl, dl, *[np.float32(g) for g in GRID], u_right=ur, native=self.native)
        mp = seq["map"]
        k0, iv, u, v, uR, lvl, vc = o.is_in_frustum(mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"], seq["cpu_poses"][i], COS_LIMIT, None)
        nm, bi, bd, qk = o.search_by_projection(F, self.sf, iv, u, v, uR, lvl, vc, mp["desc"], None, None, TH, RATIO)
        bk, bdsc = o.bird_extract(seq["bird_imgs"][i], seq["bird_mask"], BNFEAT)
        nb = 0
        if i > 0:      # the previous frame's birdview keypoints: recomputed would double the work, so they are cached per sequence
            prev = seq["cpu_bird_cache"].get(i - 1)
            if prev is None:
                prev = o.bird_extract(seq["bird_imgs"][i - 1], seq["bird_mask"], BNFEAT)
            FB = o.Frame(bk, bdsc, np.float32(0), np.float32(0), np.float32(64.0 / BW), np.float32(48.0 / BH), native=self.native)
            nb, _, _ = o.birdview_match(prev[0], prev[1], FB, None, BIRD_WINDOW, BIRD_RATIO, True)
        seq["cpu_bird_cache"][i] = (bk, bdsc)
        return nm, nb

    def prepare(self, seq):
        if "cpu_poses" not in seq:
            seq["cpu_poses"] = [self.oracle.camera_pose(**p) for p in seq["poses"]]
            seq["cpu_bird_cache"] = {}

    def run(self, seq, frames, full=False):
        """process `frames` (indices into seq) on the arm's threads; returns seconds"""
        self.prepare(seq)
        fn = self.frame_full if full else self.frame_c2
        t0 = time.perf_counter()
        if self.pool is None:
            out = [fn(seq, i) for i in frames]
        else:
            out = list(self.pool.map(lambda i: fn(seq, i), frames))
        return time.perf_counter() - t0, out

    def warm_bird_cache(self, seq, frames):
        """birdview keypoints of every frame's predecessor (so that the timed region does each frame's own work only)"""
        self.prepare(seq)
        need = sorted({i - 1 for i in frames if i > 0} - set(seq["cpu_bird_cache"]))

        def one(i):
            seq["cpu_bird_cache"][i] = self.oracle.bird_extract(seq["bird_imgs"][i], seq["bird_mask"], BNFEAT)
        if self.pool is None:
            for i in need:
                one(i)
        else:
            list(self.pool.map(one, need))


def oracle_stage_ms(img):
    """per-stage wall ms of the oracle port's extraction of one image (one thread)"""
    import oracle
    try:
        ex = oracle.Extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, native=True)
    except Exception:
        ex = oracle.Extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
    ex(img)
    ex.stage_ms(True)
    for _ in range(3):
        ex(img)
    return {k: round(v / 3, 3) for k, v in ex.stage_ms().items()}


def cv2_primitive_times(img):
    """SURVEY.md 8(d) cross-check of the scalar CPU code against OpenCV's own SIMD builds of the three stencil stages
    (cv2.resize, per-cell cv2.FAST with the 20/7 fallback, cv2.GaussianBlur), one thread, one image.  The per-cell FAST calls
    are made from Python, so the cost of the same number of calls on a 7x7 cell (no interior pixel to test) is measured and
    subtracted.  Octree, orientation and descriptors have no cv2 counterpart."""
    try:
        import cv2
    except Exception as e:
        return {"available": False, "why": repr(e)}
    cv2.setNumThreads(1)
    f20, f7 = cv2.FastFeatureDetector_create(20, True), cv2.FastFeatureDetector_create(7, True)
    tiny = np.zeros((7, 7), np.uint8)

    def once():
        t0 = time.perf_counter()
        levels, scale = [img], np.float32(1.0)
        for _ in range(1, 8):
            scale = np.float32(scale * np.float32(1.2))
            inv = np.float32(1.0) / scale
            w, h = int(round(float(np.float32(img.shape[1]) * inv))), int(round(float(np.float32(img.shape[0]) * inv)))
            levels.append(cv2.resize(levels[-1], (w, h), interpolation=cv2.INTER_LINEAR))
        t1 = time.perf_counter()
        calls = 0
        for L in levels:
            h, w = L.shape
            W_, H_ = w - 32, h - 32
            nc, nr = W_ // 30, H_ // 30
            wc, hc = -(-W_ // nc), -(-H_ // nr)
            for i in range(nr):
                y0 = 16 + i * hc
                if y0 >= h - 16 - 3:
                    continue
                y1 = min(y0 + hc + 6, h - 16)
                for j in range(nc):
                    x0 = 16 + j * wc
                    if x0 >= w - 16 - 6:
                        continue
                    c = L[y0:y1, x0:min(x0 + wc + 6, w - 16)]
                    calls += 1
                    if not f20.detect(c):
                        f7.detect(c)
                        calls += 1
        t2 = time.perf_counter()
        for L in levels:
            cv2.GaussianBlur(L, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        t3 = time.perf_counter()
        for _ in range(calls):
            f20.detect(tiny)
        t4 = time.perf_counter()
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3

    once()
    r = min((once() for _ in range(3)), key=lambda x: x[1])
    return {"available": True, "version": cv2.__version__, "threads": 1, "pyramid": round(r[0], 3), "fast": round(max(r[1] - r[3], 0.0), 3),
            "fast_call_overhead_subtracted": round(r[3], 3), "blur": round(r[2], 3)}


def cpu_stage_report(img, measured_ms_per_image):
    """Per-stage CPU denominators SURVEY.md 8(d) asks for: the oracle port's own stages, cv2's SIMD stencil stages, and the
    per-image time with every stencil stage at the faster of the two (what an OpenCV-linked build of the reference would
    approach).  `measured_ms_per_image` is what the timed CPU code actually took per image in the all-core run."""
    try:
        o = oracle_stage_ms(img)
        c = cv2_primitive_times(img)
        best = dict(o)
        if c.get("available"):
            for k in ("pyramid", "fast", "blur"):
                best[k] = min(o[k], c[k])
        so, sb = sum(o.values()), sum(best.values())
        return {"oracle_ms": o, "cv2_ms": c, "min_of_both_ms": {k: round(v, 3) for k, v in best.items()},
                "oracle_total_ms_per_image": round(so, 3), "min_of_both_total_ms_per_image": round(sb, 3),
                "speedup_of_an_opencv_simd_build_over_the_timed_code": round(so / max(sb, 1e-9), 3),
                "note": "one thread, one 1241x376 image; cv2 covers pyramid + FAST + blur only"}
    except Exception as e:
        return {"error": repr(e)}


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every few ms from a thread
    (nvidia-smi -lms cannot sample a sub-second region); falls back to one nvidia-smi query."""

    def __init__(self, device):
        self.device, self.samples, self.reasons, self.smax = device, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self._marks = []

    def _uuid_index(self):
        try:
            import torch
            return str(torch.cuda.get_device_properties(self.device).uuid)
        except Exception:
            return None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            uuid = self._uuid_index()
            if uuid:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                    u = pynvml.nvmlDeviceGetUUID(hh)
                    u = u.decode() if isinstance(u, bytes) else u
                    if uuid in u:
                        h = hh
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self._stop.is_set():
                    try:
                        self.samples.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))))
                        r = get_reasons(h)
                        for k, b in bits.items():
                            if r & b:
                                self.reasons.add(k)
                    except Exception:
                        pass
                    time.sleep(0.004)

            self._thr = threading.Thread(target=loop, daemon=True)
            self._thr.start()
        except Exception:
            self._thr = None

    def mark(self):
        self._marks.append(time.perf_counter())

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        vals = [v for t, v in self.samples if len(self._marks) < 2 or self._marks[0] <= t <= self._marks[-1]] or [v for _, v in self.samples]
        if vals:
            return {"sm_mhz": float(np.median(vals)), "sm_max_mhz": self.smax, "reasons": sorted(self.reasons), "samples": len(vals),
                    "how": "NVML polled at ~250 Hz during the timed region"}
        try:
            o = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=10).stdout.split(",")
            return {"sm_mhz": float(o[0]), "sm_max_mhz": float(o[1]), "reasons": [], "samples": 1, "how": "nvidia-smi after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}


class GpuArm:
    """P input batches (sequences of B frames with their local maps) x n_ctx contexts; steps through the frame-step C ABI."""

    def __init__(self, device, frames_per_step, pools, n_ctx=1, with_bird=True):
        import torch

        import orb_slam_birdview_b200 as pkg
        self.torch, self.pkg = torch, pkg
        self.device, self.B, self.P, self.with_bird = device, frames_per_step, pools, with_bird
        torch.cuda.set_device(device)
        self.ctxs = [pkg.Context(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, W, H, 2 * frames_per_step, device) for _ in range(n_ctx)]
        self.L = self.ctxs[0]._L
        self.cap = self.ctxs[0].max_keypoints
        self.bcap = self.L.orbb200_bird_max_keypoints(self.ctxs[0]._h, BW, BH, BNFEAT) if with_bird else 0
        self.streams = [torch.cuda.ExternalStream(self.L.orbb200_stream(c._h), device=device) for c in self.ctxs]
        self._keep = []

    # -- data ------------------------------------------------------------------------------------------
    def setup_data(self, seed0):
        torch, pkg = self.torch, self.pkg
        B, P = self.B, self.P
        dev = f"cuda:{self.device}"
        ex = pkg.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, max_size=(W, H), device=self.device)
        self.seqs, self.h_in, self.d_in, self.d_out, self.h_out, self.maps = [], [], [], [], [], []
        for p in range(P):
            seq = make_pool(B, seed0 + 1000 * p, lambda im: ex(im))
            self.seqs.append(seq)
            poses = (pkg.CameraPose * B)(*[pkg.CameraPose.make(**q) for q in seq["poses"]])
            hp = torch.from_numpy(np.frombuffer(bytes(poses), np.uint8).copy()).pin_memory()
            hi = torch.from_numpy(seq["imgs"]).pin_memory()
            hb = torch.from_numpy(seq["bird_imgs"]).pin_memory()
            self.h_in.append(dict(imgs=hi, bird=hb, poses=hp))
            self.d_in.append(dict(imgs=hi.to(dev), bird=hb.to(dev), poses=hp.to(dev)))
            i32 = dict(dtype=torch.int32, device=dev)
            self.d_out.append(dict(bi=torch.empty((B, NMAP), **i32), bd=torch.empty((B, NMAP), **i32), nm=torch.empty(B, **i32),
                                   m12=torch.empty((B, max(self.bcap, 1)), **i32), bnm=torch.empty(B, **i32)))
            mp = seq["map"]
            self.maps.append([pkg.LocalMap(c, mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"], mp["desc"]) for c in self.ctxs])
        del ex
        if self.with_bird:
            m = self.seqs[0]["bird_mask"]
            for c in self.ctxs:
                c.check(self.L.orbb200_bird_set_mask(c._h, BW, BH, BNFEAT, B, C.c_void_p(m.ctypes.data), m.strides[0]), "bird_set_mask")
        # host result buffers, one set per context (e2e)
        for _ in self.ctxs:
            u8 = lambda *s: torch.empty(s, dtype=torch.uint8).pin_memory()          # noqa: E731
            i4 = lambda *s: torch.empty(s, dtype=torch.int32).pin_memory()          # noqa: E731
            f4 = lambda *s: torch.empty(s, dtype=torch.float32).pin_memory()        # noqa: E731
            self.h_out.append(dict(kps=u8(2 * B, self.cap, 28), desc=u8(2 * B, self.cap, 32), counts=i4(2 * B), u_right=f4(B, self.cap), depth=f4(B, self.cap),
                                   bi=i4(B, NMAP), bd=i4(B, NMAP), nm=i4(B), bkps=u8(B, max(self.bcap, 1), 28), bdesc=u8(B, max(self.bcap, 1), 32),
                                   bcounts=i4(B), m12=i4(B, max(self.bcap, 1)), bnm=i4(B)))
        torch.cuda.synchronize()
        self._structs = {}

    def _params(self, full, p, ci):
        P = self.pkg.FrameStepParams()
        P.n_frames, P.w, P.h, P.stride = self.B, W, H, W
        P.min_x, P.min_y, P.inv_w, P.inv_h = GRID
        P.map = self.maps[p][ci]._h
        P.viewing_cos_limit, P.th, P.nnratio = COS_LIMIT, TH, RATIO
        if full:
            P.mb, P.mbf = MB, MBF
            P.bird_w, P.bird_h, P.bird_stride, P.bird_nfeatures = BW, BH, BW, BNFEAT
            P.bird_window, P.bird_nnratio, P.bird_check_ori = BIRD_WINDOW, BIRD_RATIO, 1
        P.chain = 0
        return P

    def _io(self, full, p, ci, host):
        """(params, inputs, outputs) ctypes structs of one (workload, input batch, context, host/device) combination, cached"""
        key = (full, p, ci, host)
        if key in self._structs:
            return self._structs[key]
        pk = self.pkg
        P, I, O = self._params(full, p, ci), pk.FrameStepInputs(), pk.FrameStepOutputs()
        src = self.h_in[p] if host else self.d_in[p]
        I.imgs, I.poses = src["imgs"].data_ptr(), src["poses"].data_ptr()
        if full:
            I.bird_imgs = src["bird"].data_ptr()
        if host:
            o = self.h_out[ci]
            O.kps, O.desc, O.counts = o["kps"].data_ptr(), o["desc"].data_ptr(), o["counts"].data_ptr()
            O.map_best_idx, O.map_best_dist, O.map_nmatches = o["bi"].data_ptr(), o["bd"].data_ptr(), o["nm"].data_ptr()
            O.cap, O.bird_cap = self.cap, self.bcap
            if full:
                O.u_right, O.depth = o["u_right"].data_ptr(), o["depth"].data_ptr()
                O.bird_kps, O.bird_desc, O.bird_counts = o["bkps"].data_ptr(), o["bdesc"].data_ptr(), o["bcounts"].data_ptr()
                O.bird_matches12, O.bird_nmatches = o["m12"].data_ptr(), o["bnm"].data_ptr()
        else:
            o = self.d_out[p]
            O.map_best_idx, O.map_best_dist, O.map_nmatches = o["bi"].data_ptr(), o["bd"].data_ptr(), o["nm"].data_ptr()
            if full:
                O.bird_matches12, O.bird_nmatches = o["m12"].data_ptr(), o["bnm"].data_ptr()
        self._structs[key] = (P, I, O)
        return self._structs[key]

    # -- steps -----------------------------------------------------------------------------------------
    def step_device(self, full, p, ci=0):
        P, I, O = self._io(full, p, ci, False)
        ctx = self.ctxs[ci]
        ctx.check(self.L.orbb200_frame_step_device(ctx._h, C.byref(P), C.byref(I), C.byref(O)), "frame_step_device")

    def step_host(self, full, p, ci=0, sub=None):
        """one host-buffer call; sub = (s, S): frames [s*B/S, (s+1)*B/S) of the batch only (a step fed as S smaller calls)"""
        P, I, O = self._io(full, p, ci, True) if sub is None else self._io_sub(full, p, ci, *sub)
        ctx = self.ctxs[ci]
        ctx.check(self.L.orbb200_frame_step_host(ctx._h, C.byref(P), C.byref(I), C.byref(O)), "frame_step_host")

    def _io_sub(self, full, p, ci, s, S):
        """structs of sub-batch s of S: the same host buffers as the whole-batch call, offset to the sub-batch's frames"""
        key = (full, p, ci, "sub", s, S)
        if key in self._structs:
            return self._structs[key]
        assert self.B % S == 0
        pk = self.pkg
        n = self.B // S
        f0 = s * n
        P, I, O = self._params(full, p, ci), pk.FrameStepInputs(), pk.FrameStepOutputs()
        P.n_frames = n
        src, o = self.h_in[p], self.h_out[ci]
        I.imgs = src["imgs"][2 * f0:].data_ptr()
        I.poses = src["poses"].data_ptr() + f0 * C.sizeof(pk.CameraPose)
        O.kps, O.desc, O.counts = o["kps"][2 * f0:].data_ptr(), o["desc"][2 * f0:].data_ptr(), o["counts"][2 * f0:].data_ptr()
        O.map_best_idx, O.map_best_dist, O.map_nmatches = o["bi"][f0:].data_ptr(), o["bd"][f0:].data_ptr(), o["nm"][f0:].data_ptr()
        O.cap, O.bird_cap = self.cap, self.bcap
        if full:
            # the sub-batches of a step run on ONE context in order: frame 0 of sub-batch s > 0 is matched against the last birdview
            # frame of sub-batch s - 1 (chain), so the step does the same work as one 128-frame call
            P.chain = 1 if s > 0 else 0
            I.bird_imgs = src["bird"][f0:].data_ptr()
            O.u_right, O.depth = o["u_right"][f0:].data_ptr(), o["depth"][f0:].data_ptr()
            O.bird_kps, O.bird_desc, O.bird_counts = o["bkps"][f0:].data_ptr(), o["bdesc"][f0:].data_ptr(), o["bcounts"][f0:].data_ptr()
            O.bird_matches12, O.bird_nmatches = o["m12"][f0:].data_ptr(), o["bnm"][f0:].data_ptr()
        self._structs[key] = (P, I, O)
        return self._structs[key]

    def h2d_bytes(self, full):
        i = self.h_in[0]
        return int(i["imgs"].numel() + i["poses"].numel() + (i["bird"].numel() if full else 0))

    def d2h_bytes(self, full):
        o = self.h_out[0]
        keys = ["kps", "desc", "counts", "bi", "bd", "nm"] + (["u_right", "depth", "bkps", "bdesc", "bcounts", "m12", "bnm"] if full else [])
        return int(sum(o[k].numel() * o[k].element_size() for k in keys))

    def launches(self):
        return sum(c.launches for c in self.ctxs)

    def sync(self):
        for c in self.ctxs:
            c.sync()

    def check_status(self):
        st = C.c_int()
        for c in self.ctxs:
            c.check(self.L.orbb200_device_status(c._h, C.byref(st)), "device_status")


def parity_check(n_frames=3, device=0):
    """The batched device step the bench times (both workloads, device and host variants) against the oracle, frame by frame."""
    import oracle
    arm = GpuArm(device, n_frames, 1, n_ctx=1)
    arm.setup_data(4242)
    seq = arm.seqs[0]
    res = {"ok": True, "frames": []}
    cpu = CpuArm(1, use_ref=False)
    cpu.prepare(seq)
    for full in (False, True):
        arm.step_device(full, 0)
        arm.sync()
        d = {k: v.cpu().numpy() for k, v in arm.d_out[0].items()}
        arm.step_host(full, 0)
        arm.sync()
        arm.check_status()
        ho = {k: v.numpy() for k, v in arm.h_out[0].items()}
        exl, exr = oracle.Extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH), oracle.Extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
        prev = None
        for i in range(n_frames):
            kl, dl = exl(seq["imgs"][2 * i])
            kr, dr = exr(seq["imgs"][2 * i + 1])
            ur = None
            if full:
                _, ur, dep = oracle.compute_stereo_matches(exl, exr, kl, dl, kr, dr, MB, MBF)
            F = oracle.Frame(kl, dl, *[np.float32(g) for g in GRID], u_right=ur)
            mp = seq["map"]
            k0, iv, u, v, uR, lvl, vc = oracle.is_in_frustum(mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"], seq["cpu_poses"][i], COS_LIMIT, None)
            n0, bi0, bd0, _ = oracle.search_by_projection(F, exl.scale_factors(), iv, u, v, uR, lvl, vc, mp["desc"], None, None, TH, RATIO)
            cnt = ho["counts"]
            kps_l = ho["kps"][2 * i].view(arm.pkg.KP_DTYPE).reshape(-1)[:cnt[2 * i]]
            kps_r = ho["kps"][2 * i + 1].view(arm.pkg.KP_DTYPE).reshape(-1)[:cnt[2 * i + 1]]
            ok = (int(d["nm"][i]) == n0 and np.array_equal(d["bi"][i], bi0) and np.array_equal(d["bd"][i][bi0 >= 0], bd0[bi0 >= 0])
                  and kps_l.tobytes() == kl.tobytes() and kps_r.tobytes() == kr.tobytes()
                  and np.array_equal(ho["desc"][2 * i][:len(kl)], dl) and np.array_equal(ho["desc"][2 * i + 1][:len(kr)], dr)
                  and np.array_equal(ho["bi"][i], bi0) and int(ho["nm"][i]) == n0)
            nb0 = 0
            if full:
                bk, bdsc = oracle.bird_extract(seq["bird_imgs"][i], seq["bird_mask"], BNFEAT)
                nbk = int(ho["bcounts"][i])
                ok = ok and np.array_equal(ho["u_right"][i][:len(kl)].view(np.uint32), ur.view(np.uint32))
                ok = ok and ho["bkps"][i].view(arm.pkg.KP_DTYPE).reshape(-1)[:nbk].tobytes() == bk.tobytes() and np.array_equal(ho["bdesc"][i][:nbk], bdsc)
                if prev is not None:
                    FB = oracle.Frame(bk, bdsc, np.float32(0), np.float32(0), np.float32(64.0 / BW), np.float32(48.0 / BH))
                    nb0, m12, _ = oracle.birdview_match(prev[0], prev[1], FB, None, BIRD_WINDOW, BIRD_RATIO, True)
                    ok = ok and int(ho["bnm"][i]) == nb0 and np.array_equal(ho["m12"][i][:len(m12)], m12) and int(d["bnm"][i]) == nb0
                prev = (bk, bdsc)
            res["frames"].append({"workload": "C3_full" if full else "C2", "nmatches": int(d["nm"][i]), "oracle": int(n0), "bird_matches": int(nb0), "ok": bool(ok)})
            res["ok"] = res["ok"] and bool(ok)
    cpu.close()
    return res


RANK_PLACEMENT = "packed"      # how dist_setup mapped the local ranks to GPUs (goes into the JSON line)


def place_rank(local, world, ndev):
    """Which GPU a local rank uses when fewer ranks than GPUs run on the box.  Pinned host-to-device copies are capped per group of
    four GPUs of an 8-GPU box (~116-128 GB/s for GPUs 0-3, the same for 4-7; one or two copying ranks per group lose nothing, four get
    29 GB/s each instead of 55: profiles/r2f_n4_placement.json), and the end-to-end leg is bound by exactly those copies -- so the ranks
    alternate between the two halves: 0, 4, 1, 5, ...  With as many ranks as visible GPUs (or ORBB200_RANK_PLACEMENT=packed) rank i
    uses GPU i."""
    if os.environ.get("ORBB200_RANK_PLACEMENT", "") == "packed" or world < 2 or ndev < 4 or ndev % 2 or world >= ndev:
        return local, "packed"
    return (local // 2) + (local % 2) * (ndev // 2), "spread over the two halves of the box (host-I/O groups)"


def dist_setup(args):
    global RANK_PLACEMENT
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        local, RANK_PLACEMENT = place_rank(local, world, torch.cuda.device_count() if args.impl == "ours" else 0)
        torch.cuda.set_device(local)
        # stdout carries ONE JSON line: NCCL prints its version banner (and, with NCCL_DEBUG=INFO, its log) to file descriptor 1 when
        # the communicator is created, so descriptor 1 points at stderr while that happens
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist_.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
            dist_.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        dist = dist_
    return rank, world, local, dist


_FULL_AFFINITY = None      # the process's CPU set before bind_near_gpu narrowed it (the CPU baseline gets all of it back)


def bind_near_gpu(device, enable=True):
    """Pin this process (and the threads / pinned host allocations it makes afterwards) to the CPUs NVML reports as
    closest to the GPU, so that each rank's staging buffers live on the NUMA node its GPU's PCIe root hangs off."""
    info = {"enabled": bool(enable), "cpus_before": len(os.sched_getaffinity(0))}
    if not enable:
        return info
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = None
        try:
            import torch
            uuid = "GPU-" + str(torch.cuda.get_device_properties(device).uuid)
        except Exception:
            pass
        h = None
        if uuid:
            for i in range(pynvml.nvmlDeviceGetCount()):
                hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(hh)
                if (u.decode() if isinstance(u, bytes) else u) == uuid:
                    h = hh
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(device)
        before = os.sched_getaffinity(0)
        global _FULL_AFFINITY
        _FULL_AFFINITY = set(before)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        ideal = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus = (ideal & before) or before          # stay inside the container's cpuset
        os.sched_setaffinity(0, cpus)
        info.update({"cpus_after": len(cpus), "first_cpu": min(cpus), "last_cpu": max(cpus)})
    except Exception as e:
        info["error"] = repr(e)
    return info


# ---- timed legs ----------------------------------------------------------------------------------------------
def leg_device(arm, full, K, Wm, nctx, barrier, sampler=None):
    """K device-resident steps alternating over nctx contexts (streams): the tail of one step (matching: latency-bound CTAs)
    overlaps the head of the next.  CUDA events on the contexts' streams; returns (ms, launches)."""
    torch, P = arm.torch, arm.P
    for s in range(P * nctx):                  # pre-warm, not counted as warm-up: the first use of an (input batch, context) pair
        arm.step_device(full, s % P, s % nctx)  # builds its step plan (cudaMalloc + graph capture)
    arm.sync()
    for s in range(Wm):
        arm.step_device(full, s % P, s % nctx)
    arm.sync()
    barrier()
    l0 = arm.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.mark()
    e0.record(arm.streams[0])
    for i in range(1, nctx):
        arm.streams[i].wait_event(e0)
    for s in range(K):
        arm.step_device(full, s % P, s % nctx)
    for i in range(1, nctx):
        join = torch.cuda.Event()
        join.record(arm.streams[i])
        arm.streams[0].wait_event(join)
    e1.record(arm.streams[0])
    barrier()
    if sampler:
        sampler.mark()
    return e0.elapsed_time(e1), arm.launches() - l0


def leg_stages(arm, full, K):
    """per-stage durations: a second pass of the same K steps with events between the stages on ONE context, so that a
    stage's duration is that of its kernels running alone (the timed region replays the captured graphs)"""
    c = arm.ctxs[0]
    arm.L.orbb200_stage_timing(c._h, 1)
    arm.L.orbb200_stage_times(c._h, None, None, 1)
    for s in range(K):
        arm.step_device(full, s % arm.P, 0)
    c.sync()
    ms, n = np.zeros(NSTAGES, np.float32), np.zeros(NSTAGES, np.int32)
    arm.L.orbb200_stage_times(c._h, C.c_void_p(ms.ctypes.data), C.c_void_p(n.ctypes.data), 1)
    arm.L.orbb200_stage_timing(c._h, 0)
    return ms, n


def leg_e2e(arm, full, K, Wm, NE, barrier, S=1):
    """K steps through the host-buffer C-ABI call, NE contexts in flight; wall clock around the region; returns seconds.
    S > 1: a step's batch is handed over as S calls of B/S frames each (what a caller does to overlap its uploads with the
    device work from the first frame on: the un-overlapped first upload and last download of the timed region shrink S-fold).
    C2: the calls rotate over the contexts.  North-star frame: the S calls of a step stay on one context, in order, chained
    (the birdview match of a call's first frame needs the previous call's last frame); the contexts rotate per step."""
    torch, P = arm.torch, arm.P
    sub = (lambda s: None) if S == 1 else (lambda s: (s, S))
    ctx_of = (lambda j: (j // S) % NE) if full else (lambda j: j % NE)
    for j in range(max(P * NE, 2 * NE) * S):   # pre-warm: plans + staging buffers of every (input batch, context) pair
        arm.step_host(full, (j // S) % P, ctx_of(j), sub(j % S))
    arm.sync()
    for j in range(Wm * S):
        arm.step_host(full, (j // S) % P, ctx_of(j), sub(j % S))
    arm.sync()
    barrier()
    reuse = NE * S if full else NE          # calls between two uses of the same host result buffers
    t0 = time.perf_counter()
    for j in range(K * S):
        ci = ctx_of(j)
        if j >= reuse and (not full or j % S == 0):
            arm.ctxs[ci].sync()          # results of the context's previous step / call are on the host: consume before reuse
            _ = int(arm.h_out[ci]["nm"][0])
        arm.step_host(full, (j // S) % P, ci, sub(j % S))
    arm.sync()
    _ = int(arm.h_out[0]["nm"][0])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    arm.check_status()
    return dt


def h2d_rate(arm, barrier=None):
    """pinned host->device copy rate of this rank in GB/s (one stream, the C2 image batch)"""
    torch = arm.torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    src, dst = arm.h_in[0]["imgs"], arm.d_in[0]["imgs"]
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if barrier:
        barrier()
    e0.record()
    for _ in range(4):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return 4 * src.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9


def host_memcpy_gbs(seconds=0.5):
    """plain memcpy rate over all host cores (numpy copies, one 256 MB buffer per thread): the host DRAM ceiling of the box"""
    from concurrent.futures import ThreadPoolExecutor
    n = host_cores()
    bufs = [(np.ones(1 << 26, np.float32), np.empty(1 << 26, np.float32)) for _ in range(min(n, 16))]

    def work(ab):
        a, b = ab
        t_end, moved = time.perf_counter() + seconds, 0
        while time.perf_counter() < t_end:
            np.copyto(b, a)
            moved += 2 * a.nbytes
        return moved
    t0 = time.perf_counter()
    with ThreadPoolExecutor(len(bufs)) as pool:
        tot = sum(pool.map(work, bufs))
    return tot / (time.perf_counter() - t0) / 1e9, len(bufs)


# ---- secondary configs -------------------------------------------------------------------------------------------
def ev_time(torch, stream, fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def leg_extract_only(torch, pkg, device, name, w, h, nfeat, batch, reps=10):
    """C1 / C5: extraction only, device-resident batch; plus one image per call through the host API"""
    synth = _synth()
    L = pkg.load_library()
    ctx = pkg.Context(nfeat, SCALE, NLEVELS, INI_TH, MIN_TH, w, h, batch, device)
    stream = torch.cuda.ExternalStream(L.orbb200_stream(ctx._h), device=device)
    base = [synth.synth_frame(h, w, 7000 + i) for i in range(min(batch, 6))]
    imgs = np.stack([base[i % len(base)] if i < len(base) else synth.shift_frame(base[i % len(base)], i // len(base), 0) for i in range(batch)])
    d = torch.from_numpy(imgs).to(f"cuda:{device}")
    ms = ev_time(torch, stream, lambda: ctx.check(L.orbb200_extract_device(ctx._h, d.data_ptr(), w * h, batch, w, h, w)), reps)
    ex = pkg.ORBextractor(nfeat, SCALE, NLEVELS, INI_TH, MIN_TH, max_size=(w, h), device=device)
    ex(imgs[0])
    t0 = time.perf_counter()
    for _ in range(20):
        k, _d = ex(imgs[0])
    lat = (time.perf_counter() - t0) / 20
    sum_p, p0 = level_pixels(w, h)
    return {"workload": f"{name}: ORBextractor {nfeat} feats on {w}x{h}, extraction only", "batch": batch, "value": batch / (ms * 1e-3), "unit": "frames/s",
            "ms_per_batch": ms, "one_image_per_call_host_api_ms": lat * 1e3, "keypoints": int(len(k)),
            "algorithmic_bytes_per_frame": int(3 * sum_p - p0 + 60 * len(k)),
            "l2": f"{batch}-image pyramid + blur pools ({2 * batch * sum_p * 1.05 / 1e6:.0f} MB) > 126 MB L2" if 2 * batch * sum_p > 126e6 else "pools fit L2: re-reads may hit it"}


def leg_knn2(torch, pkg, device, reps=10):
    """C4: brute-force Hamming best / second best, 2k queries vs 2k / 20k / 200k map descriptors, device-resident"""
    synth = _synth()
    L = pkg.load_library()
    ctx = pkg.Context(1000, SCALE, NLEVELS, INI_TH, MIN_TH, 64, 64, 1, device)
    stream = torch.cuda.ExternalStream(L.orbb200_stream(ctx._h), device=device)
    peak = float(L.orbb200_measure_popc_peak(ctx._h))       # G popc/s measured on this device
    out = {"popc_peak_gops_measured": peak, "sizes": []}
    nq = 2000
    q = torch.from_numpy(synth.synth_descriptors(nq, 1)).to(f"cuda:{device}")
    for nm in (2000, 20000, 200000):
        m = torch.from_numpy(synth.synth_descriptors(nm, 2)).to(f"cuda:{device}")
        bi = torch.empty(nq, dtype=torch.int32, device=f"cuda:{device}")
        bd, sd = torch.empty_like(bi), torch.empty_like(bi)
        ms = ev_time(torch, stream, lambda: ctx.check(L.orbb200_hamming_knn2_device(ctx._h, q.data_ptr(), nq, m.data_ptr(), nm, bi.data_ptr(),
                                                                                  bd.data_ptr(), sd.data_ptr())), reps)
        pairs = nq * nm / (ms * 1e-3)
        out["sizes"].append({"nq": nq, "nm": nm, "ms": ms, "gpairs_per_s": pairs * 1e-9, "popc_frac": pairs * 8 * 1e-9 / peak if peak > 0 else None})
    return out


def leg_per_frame(arm, frames=12):
    """The drop-in path: ONE frame per call through the host API, every result back on the host before the next call --
    ORBextractor on L and R, ComputeStereoMatches, SearchLocalPoints, the birdview front-end and SearchByMatchBird, exactly
    the calls Frame::Frame + Tracking::Track make.  Wall ms per frame."""
    pkg, seq = arm.pkg, arm.seqs[0]
    ctx = pkg.Context(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, W, H, 2, arm.device)
    mp = seq["map"]
    M = pkg.LocalMap(ctx, mp["pos"], mp["normal"], mp["max_distance"], mp["min_distance"], mp["desc"])
    step = pkg.FrameStep(ctx, W, H, M, mb=MB, mbf=MBF, th=TH, nnratio=RATIO, bird_size=(BW, BH), bird_nfeatures=BNFEAT,
                         bird_mask=seq["bird_mask"], bird_window=BIRD_WINDOW, bird_nnratio=BIRD_RATIO)
    poses = [pkg.CameraPose.make(**p) for p in seq["poses"]]
    n = min(frames, len(poses))
    step(seq["imgs"][0:2], seq["bird_imgs"][0:1], poses[0:1])
    t0 = time.perf_counter()
    for i in range(1, n):
        step(seq["imgs"][2 * i:2 * i + 2], seq["bird_imgs"][i:i + 1], poses[i:i + 1], chain=True)
    full_ms = (time.perf_counter() - t0) / (n - 1) * 1e3
    ex = pkg.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, max_size=(W, H), device=arm.device)
    ex(seq["imgs"][0])
    t0 = time.perf_counter()
    for i in range(n):
        ex(seq["imgs"][2 * i])
    ex_ms = (time.perf_counter() - t0) / n * 1e3
    out = {"what": "one frame per host call, results on the host before the next call (the reference's call pattern)",
           "north_star_frame_ms": full_ms, "north_star_frames_per_s": 1e3 / full_ms, "one_1241x376_extraction_ms": ex_ms}
    # the same through the C++ drop-in class (cpp/ORBextractor.h::operator(), as Frame::ExtractORB calls it)
    try:
        import tempfile
        drv = os.path.join(ROOT, "orb-slam-birdview_b200", "cpp", "shim_driver")
        if not os.path.exists(drv):
            subprocess.run(["make", "-s", "-C", os.path.dirname(drv), "shim_driver"], check=True, timeout=300)
        shim = {}
        with tempfile.TemporaryDirectory() as d:
            for (h, w, nf) in ((480, 752, 1000), (H, W, NFEAT)):
                img = seq["imgs"][0] if (h, w) == (H, W) else _synth().synth_frame(h, w, 1000)
                raw = os.path.join(d, "in.raw")
                np.ascontiguousarray(img).tofile(raw)
                r = subprocess.run([drv, raw, str(w), str(h), str(nf), str(INI_TH), str(MIN_TH), os.path.join(d, "out.bin")], capture_output=True, text=True,
                                   timeout=300, env=dict(os.environ, ORBB200_SHIM_TIME="200"))
                line = [ln for ln in r.stdout.splitlines() if ln.startswith("TIMING ")]
                if line:
                    shim[f"{w}x{h}"] = json.loads(line[0][7:])
        if shim:
            out["cpp_shim"] = shim
        # the same frame step and the same extraction from C++ with no interpreter in the loop (tools/ubench/*.cu, built by
        # __graft_entry__.build()): wall ms per call and the device chain alone (CUDA events)
        cpp = {}
        for name, args in (("frame_timeline", ["100"]), ("call_timeline", ["752", "480", "1000", "200"]), ("call_timeline", ["1241", "376", "2000", "200"]),
                           ("concurrent_calls", ["1241", "376", "2000", "200"])):
            exe = os.path.join(ROOT, "tools", "ubench", name)
            if os.path.exists(exe):
                r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
                line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                if line:
                    cpp[name if name == "frame_timeline" else f"{name}_{args[0]}x{args[1]}"] = json.loads(line[-1])
        if cpp:
            out["cpp_c_abi"] = cpp
    except Exception as e:      # the C++ number is a side leg: never fail the bench line for it
        out["cpp_shim"] = {"error": repr(e)}
    return out


def c5_sharded_digest(torch, pkg, dist, rank, world, device, n_frames=256, w=1920, h=1080, nfeat=4000):
    """C5 determinism check: a fixed 256-frame 1080p sequence split over the ranks with a 1-frame overlap; per
    frame: extraction + brute-force matching against the previous frame.  The SHA-1 over all frames' (keypoints, descriptors,
    matches) must not depend on the number of ranks.  Returns (digest, frames/s over all ranks, per-rank frames)."""
    synth = _synth()
    from importlib import import_module
    shard = import_module("orb_slam_birdview_b200.shard")
    step_px = 3
    canvas = synth.synth_frame(h, w + step_px * n_frames + 8, 515151)
    read0, own0, end = shard.shard_ranges(n_frames, world, overlap=1)[rank]
    L = pkg.load_library()
    batch = 16
    ctx = pkg.Context(nfeat, SCALE, NLEVELS, INI_TH, MIN_TH, w, h, batch, device)
    cap = ctx.max_keypoints
    dev = f"cuda:{device}"
    digests, prev = {}, None
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for b0 in range(read0, end, batch):
        idx = list(range(b0, min(b0 + batch, end)))
        imgs = np.stack([canvas[:, step_px * i:step_px * i + w] for i in idx])
        d = torch.from_numpy(np.ascontiguousarray(imgs)).to(dev)
        ctx.check(L.orbb200_extract_device(ctx._h, d.data_ptr(), w * h, len(idx), w, h, w), "extract_device")
        k = np.empty((len(idx), cap), pkg.KP_DTYPE)
        ds = np.empty((len(idx), cap, 32), np.uint8)
        n = np.empty(len(idx), np.int32)
        ctx.check(L.orbb200_download_results(ctx._h, len(idx), C.c_void_p(k.ctypes.data), C.c_void_p(ds.ctypes.data), cap, C.c_void_p(n.ctypes.data)), "download")
        for j, i in enumerate(idx):
            kj, dj = k[j][:n[j]], ds[j][:n[j]]
            if i >= own0:
                hsh = hashlib.sha1()
                hsh.update(kj.tobytes())
                hsh.update(dj.tobytes())
                if prev is not None and len(prev) and len(dj):
                    for a in pkg.ORBmatcher(ctx).hamming_knn2(dj, prev):
                        hsh.update(a.tobytes())
                digests[i] = hsh.digest()
            prev = dj
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    mine = np.zeros((n_frames, 20), np.uint8)
    for i, dg in digests.items():
        mine[i] = np.frombuffer(dg, np.uint8)
    t = torch.from_numpy(mine).to(dev).to(torch.int32)
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)         # every frame is owned by exactly one rank
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    allh = hashlib.sha1(t.cpu().numpy().astype(np.uint8).tobytes()).hexdigest()
    return allh, n_frames / float(tt[0]), end - own0


# ---- our arm -----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    rank, world, local, dist = dist_setup(args)
    host_bind = bind_near_gpu(local, not args.no_numa_bind)
    B, P = args.frames_per_step, args.pools
    NE = max(2, args.e2e_contexts)
    SUB = args.e2e_calls_per_step if args.e2e_calls_per_step > 0 and B % max(args.e2e_calls_per_step, 1) == 0 else 1
    arm = GpuArm(local, B, P, n_ctx=NE, with_bird=not args.headline_only)
    arm.setup_data(100000 * rank + 2000)
    K, Wm = args.steps, max(args.warmup, 3)
    dev = f"cuda:{local}"

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(*vals):
        if dist is None:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    nctx = max(1, min(args.device_contexts, NE))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- headline C2: device-resident (value), stages, end to end ----
    ms, launches = leg_device(arm, False, K, Wm, nctx, barrier, sampler)
    clocks = sampler.stop() if rank == 0 else None
    st_ms, st_n = leg_stages(arm, False, K)
    e2e_solo_s = None
    if world > 1 and not args.no_e2e:
        # the same leg with rank 0 alone on the host side: numerator of the end-to-end scaling efficiency
        if rank == 0:
            e2e_solo_s = leg_e2e(arm, False, K, Wm, NE, lambda: torch.cuda.synchronize(), SUB)
        barrier()
    e2e_s = leg_e2e(arm, False, K, Wm, NE, barrier, SUB) if not args.no_e2e else float("nan")
    h2d_conc = h2d_rate(arm, barrier)
    # ---- C3_full: the north-star frame ----
    full = None
    if not args.headline_only:
        f_ms, f_launches = leg_device(arm, True, K, Wm, nctx, barrier)
        f_st_ms, f_st_n = leg_stages(arm, True, K)
        f_e2e_s = leg_e2e(arm, True, K, Wm, NE, barrier, 1) if not args.no_e2e else float("nan")   # (4 chained calls per step measured slower: 6.94 vs 6.32 ms)
        f_ms, f_e2e_ms = maxr(f_ms, f_e2e_s * 1e3)
        full = dict(ms=f_ms, e2e_ms=f_e2e_ms, launches=f_launches, st_ms=f_st_ms, st_n=f_st_n)
    # ---- C5 sharded determinism digest (all ranks) ----
    c5 = None
    if not args.headline_only and not args.no_side_configs:
        try:
            c5 = c5_sharded_digest(torch, arm.pkg, dist, rank, world, local)
        except Exception as e:          # reported, never fatal for the headline
            c5 = ("failed: " + repr(e), None, 0)
    ms, e2e_ms, neg_h2d = maxr(ms, e2e_s * 1e3, -h2d_conc)
    h2d_conc_min, h2d_conc_sum = -neg_h2d, h2d_conc
    if dist is not None:
        t2 = torch.tensor([h2d_conc], device=dev, dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.SUM)
        h2d_conc_sum = float(t2[0])
    if rank != 0:
        if dist is not None:
            dist.barrier()               # rank 0's side measurements
            dist.destroy_process_group()
        return

    h2d_gbs = h2d_rate(arm)              # this rank alone
    frames_total = world * B * K
    value = frames_total / (ms * 1e-3)
    e2e_value = frames_total / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel of the headline step ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    avg = st_ms / np.maximum(st_n, 1)
    dom = int(np.argmax(avg))
    sum_p, p0 = level_pixels(W, H)
    # algorithmic bytes per launch group: resize l: P_{l-1}+P_l ; blur l: 2 P_l ; FAST l: P_l
    alg = {"pyramid": 2 * sum_p - p0 - level_pixels(W, H, last=True), "fast": sum_p, "blur": 2 * sum_p, "import": 2 * p0}
    name = STAGES[dom]
    bytes_per_launch = alg.get(name)
    traffic, ncu_pipes, t = None, {}, None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json"))).get(name)
        traffic = t["dram_bytes_per_image"] * 2 * B if t else None   # ncu dram read+write, scaled to this launch's image count
        ncu_pipes = {k: t[k] for k in ("alu_pipe_pct_of_peak", "issue_active_pct_of_peak", "lsu_pipe_pct_of_peak", "l1tex_throughput_pct_of_peak") if t and k in t}
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": name, "unit": "GB/s", "peak": hbm_peak, "peak_source": peak_src, "traffic": traffic,
                "avg_launch_ms": float(avg[dom])}
    if bytes_per_launch is not None:
        per_launch = bytes_per_launch * 2 * B
        roofline["achieved"] = per_launch / (float(avg[dom]) * 1e-3) * 1e-9
        roofline["frac"] = roofline["achieved"] / hbm_peak
        roofline["algorithmic_bytes_per_launch"] = int(per_launch)
        if ncu_pipes:
            roofline["ncu_pipe_utilisation"] = dict(ncu_pipes, source=str(t.get("source", "profiles/")))
    else:
        roofline["achieved"] = None
        roofline["frac"] = None
    roofline["note"] = "FAST arc tests are integer-ALU bound long before HBM (no tensor-core work on this path); frac is vs the HBM copy peak"

    configs = {}
    if full is not None:
        configs["C3_full"] = {
            "workload": "north-star frame: C2 + ComputeStereoMatches + birdview 400x400 cv::ORB(2000) detect(mask) + cornerSubPix + compute + "
                        "SearchByMatchBird(window 15) against the previous frame",
            "value": frames_total / (full["ms"] * 1e-3), "unit": "frames/s", "ms_per_step": full["ms"] / K,
            "e2e": {"value": frames_total / (full["e2e_ms"] * 1e-3), "unit": "frames/s", "ms_per_step": full["e2e_ms"] / K, "calls_per_step": 1,
                    "h2d_bytes_per_step": arm.h2d_bytes(True), "d2h_bytes_per_step": arm.d2h_bytes(True)},
            "gpu_launches": int(full["launches"]),
            "stage_ms_per_step": {STAGES[i]: float(full["st_ms"][i] / K) for i in range(NSTAGES) if STAGES[i] != "-"}}
    if c5 is not None:
        configs["C5_sharded"] = {"workload": "C5: one 256-frame 1920x1080/4000 sequence split over the ranks (1-frame overlap), extraction + "
                                             "brute-force matching against the previous frame, results gathered on the host",
                                 "shard_digest": c5[0], "value": c5[1], "unit": "frames/s (host API, per-batch download + hashing inside)",
                                 "frames_rank0": int(c5[2])}
    if world == 1 and not args.headline_only and not args.no_side_configs:
        for nm_, fn in (("C1", lambda: leg_extract_only(torch, arm.pkg, local, "C1", 752, 480, 1000, 256)),
                        ("C5", lambda: leg_extract_only(torch, arm.pkg, local, "C5", 1920, 1080, 4000, 32)),
                        ("C4", lambda: leg_knn2(torch, arm.pkg, local)),
                        ("per_frame", lambda: leg_per_frame(arm))):
            try:
                configs[nm_] = fn()
            except Exception as e:
                configs[nm_] = {"error": repr(e)}

    # ---- CPU baseline on this box's host cores: bounded samples of the same workloads ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline(arm.seqs[0], full is not None)
        if full is not None and cpu.get("C3_full"):
            configs["C3_full"]["cpu_baseline"] = cpu.pop("C3_full")
            cb = configs["C3_full"]["cpu_baseline"]
            if cb.get("value"):
                configs["C3_full"]["e2e_vs_cpu_all_cores"] = configs["C3_full"]["e2e"]["value"] / cb["value"]
    host_gbs = None
    try:
        host_gbs = host_memcpy_gbs()
    except Exception:
        pass

    out = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": config_dict(B, P),
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": arm.h2d_bytes(False), "d2h_bytes_per_step": arm.d2h_bytes(False),
                "ms_per_step": e2e_ms / K, "contexts_in_flight": NE, "calls_per_step": SUB, "frames_per_call": B // SUB, "h2d_copy_gbs_measured": h2d_gbs, "host_binding": host_bind},
        "e2e_h2d_gbs_all_ranks_at_once_min_rank": h2d_conc_min, "e2e_h2d_gbs_all_ranks_at_once_sum": h2d_conc_sum,
        "e2e_scaling_efficiency": (e2e_value / world) / (B * K / e2e_solo_s) if e2e_solo_s else (1.0 if world == 1 else None),
        "e2e_solo_rank0_frames_per_s": (B * K / e2e_solo_s) if e2e_solo_s else None,
        "host_memcpy_gbs_all_cores": {"value": host_gbs[0], "threads": host_gbs[1]} if host_gbs else None,
        "gpu_launches": int(launches),
        "rank_placement": {"how": RANK_PLACEMENT, "rank0_gpu": int(local), "visible_gpus": int(torch.cuda.device_count())},
        "clocks": clocks,
        "roofline": roofline,
        "stage_ms_per_step": {STAGES[i]: float(st_ms[i] / K) for i in range(9)},
        "stage_timing": "CUDA events between the stages in a second pass of the same K steps on one context (the timed region replays the captured graphs, blur forked beside FAST)",
        "device_contexts": nctx,
        "configs": configs,
        "cpu_baseline": cpu,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(seq, with_full, target_s=12.0):
    """The reference on the host cores; bounded samples (about target_s seconds of wall time each)."""
    try:
        if _FULL_AFFINITY:
            os.sched_setaffinity(0, _FULL_AFFINITY)      # the baseline may use every host core, not only the GPU's neighbours
        cores = host_cores()
        navail = len(seq["poses"])
        one = CpuArm(1)
        t1 = min(one.run(seq, [0])[0], one.run(seq, [0])[0])          # one frame, one thread: sizes the sample
        cpu = CpuArm(cores)
        n = int(min(max(cores * 2, target_s * cores / max(t1, 1e-3)), 400 * cores))
        frames = [i % navail for i in range(n)]
        cpu.run(seq, frames[:cores])                                   # warm-up: per-thread extractors
        dt, _ = cpu.run(seq, frames)
        out = {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "reference" if cpu.use_ref else "port",
               "sample": f"{n} stereo frames of the C2 workload on {cores} host threads ({dt:.1f} s); 1 thread: {1.0 / t1:.2f} frames/s; "
                         + ("ORBextractor = the reference's own src/ORBextractor.cc compiled unmodified (oracle/_ref; its OpenCV primitives are the "
                            "cv2-pinned scalar restatements), isInFrustum + SearchByProjection = oracle port" if cpu.use_ref else "oracle port"),
               "single_thread_value": 1.0 / t1, "native_build": cpu.native, "cpu_model": host_cpu_model(),
               "per_stage": cpu_stage_report(seq["imgs"][0], 1e3 * cores / (2 * n / dt))}
        if with_full:
            t1f = one.run(seq, [1], full=True)[0]
            nf = int(min(max(cores * 2, target_s * cores / max(t1f, 1e-3)), 400 * cores))
            ff = [1 + (i % (navail - 1)) for i in range(nf)]
            cpu.warm_bird_cache(seq, ff)
            cpu.run(seq, ff[:cores], full=True)
            dtf, _ = cpu.run(seq, ff, full=True)
            out["C3_full"] = {"value": nf / dtf, "unit": "frames/s", "cores": cores, "kind": "port",
                              "sample": f"{nf} north-star frames on {cores} host threads ({dtf:.1f} s); 1 thread: {1.0 / t1f:.2f} frames/s; oracle port "
                                        "(ComputeStereoMatches reads both extractors' pyramids, so the port's extractor is used throughout)",
                              "single_thread_value": 1.0 / t1f,
                              "reference_configuration_value": 2.0 / t1f,
                              "reference_configuration": "what the reference itself does: one frame at a time, L and R extraction on two threads (src/Frame.cc:124-127) -- "
                                                         "an upper bound of 2x the one-thread figure"}
        one.close()
        cpu.close()
        return out
    except Exception as e:   # the baseline is reported, never required for the GPU number
        return {"value": None, "unit": "frames/s", "cores": 0, "kind": "port", "sample": f"failed: {e!r}"}


def run_reference(args):
    """--impl reference: the reference on all host threads, same workload / metric / config / steps.  Extraction is the
    reference's own src/ORBextractor.cc compiled unmodified (oracle/_ref); the matcher side is the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import oracle
    cores = host_cores()
    cpu = CpuArm(cores)
    B, P = args.frames_per_step, args.pools
    F = B if args.ref_frames_per_step <= 0 else args.ref_frames_per_step
    npool = min(F, 32)
    port = oracle.Extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
    seq = make_pool(npool, 2000, lambda im: port(im))
    frames = [i % npool for i in range(F)]
    K, Wm = args.steps, args.warmup
    for _ in range(max(Wm, 1)):
        cpu.run(seq, frames)
    # the K steps go to the persistent pool back to back: a step's first frames start while the previous step's last frames finish,
    # as the CUDA arm's steps overlap across its contexts (a barrier after every step cost the arm 15 % against the long-run
    # cpu_baseline of the same code: 16 threads idle for half a frame time 40 times)
    t0 = time.perf_counter()
    cpu.run(seq, frames * K)
    dt = time.perf_counter() - t0
    value = F * K / dt
    configs = {}
    if not args.headline_only:
        ff = [1 + (i % (npool - 1)) for i in range(F)]
        cpu.warm_bird_cache(seq, ff)
        cpu.run(seq, ff[:cores], full=True)
        kf = max(1, min(K, 3))
        t0 = time.perf_counter()
        for _ in range(kf):
            cpu.run(seq, ff, full=True)
        dtf = time.perf_counter() - t0
        configs["C3_full"] = {"value": F * kf / dtf, "unit": "frames/s", "steps": kf, "kind": "port", "cores": cores}
    kind = "reference" if cpu.use_ref else "port"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_dict(B, P),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "cpu_model": host_cpu_model(),
                         "sample": f"{F} stereo frames per step on {cores} host threads (persistent pool, per-thread extractors), native_build={cpu.native}; "
                                   "ORBextractor = the reference's own source compiled unmodified (oracle/_ref), matcher = oracle port"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "configs": configs,
    }
    cpu.close()
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=128)
    ap.add_argument("--pools", type=int, default=3)
    ap.add_argument("--ref-frames-per-step", type=int, default=0)
    ap.add_argument("--headline-only", action="store_true", help="C2 only: no north-star leg, no side configs (profiling runs)")
    ap.add_argument("--no-side-configs", action="store_true", help="skip C1 / C4 / C5 / per-frame legs")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the host-core baseline (profiling runs)")
    ap.add_argument("--device-contexts", type=int, default=2, help="contexts (streams) the device-resident leg alternates its steps over")
    ap.add_argument("--e2e-contexts", type=int, default=8, help="contexts (streams) the host-buffer leg keeps in flight")
    ap.add_argument("--e2e-calls-per-step", type=int, default=4, help="the C2 host-buffer leg hands a step's batch over as this many calls (frames_per_step must be a multiple)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the rank to the CPUs nearest its GPU")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
