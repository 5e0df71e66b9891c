#!/usr/bin/env python3
"""bench.py -- ORB front-end throughput on B200 (BASELINE.json metric: extract+match frames/s).

Workload (config C2, BASELINE.json configs[1]): KITTI-shape stereo frames, 1241x376 left+right, 2000
features per image, 8 levels, scale 1.2, FAST 20/7; per frame: ORBextractor on both images, lookup grid
on the left frame, SearchByProjection of 3000 local-map points (th=1, ratio 0.8).  One "step" = one
pass of that hot path over a batch of --frames-per-step synthetic stereo frames.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU)
  python bench.py --impl reference --steps K --warmup W     the reference algorithm on the host cores
                                                            (C++ oracle port, all host threads)

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events); `e2e` = the same step
through the host-buffer C-ABI call, H2D/D2H inside the timed region; `roofline` = the dominant kernel
against the measured HBM peak; `cpu_baseline` = the oracle on the box's host cores (bounded sample).
"""
import argparse
import ctypes as C
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
NQ, TH, RATIO = 3000, 1.0, 0.8
GRID = (0.0, 0.0, 64.0 / W, 48.0 / H)        # mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv (k1 == 0)
METRIC = "ORB extract+match frames/sec"
STAGES = ["import", "pyramid", "fast", "blur", "octree", "describe", "grid", "match", "stereo"]
MB, MBF = 0.537, 386.1448                      # KITTI00-02.yaml: Camera.bf / Camera.fx, Camera.bf


def _synth():
    from importlib import import_module
    return import_module("orb_slam_birdview_b200.synth")


def make_images(n_frames, seed0, w=W, h=H):
    synth = _synth()
    imgs = np.empty((2 * n_frames, h, w), np.uint8)
    for i in range(n_frames):
        left = synth.synth_frame(h, w, seed0 + i)
        imgs[2 * i] = left
        imgs[2 * i + 1] = synth.shift_frame(left, -7, 0)       # right view: 7 px disparity
    return imgs


def make_queries(kps_left, desc_left, counts_left, nq, seed0, w=W, h=H):
    """[n_frames][nq] query arrays from each left frame's own keypoints"""
    synth = _synth()
    n = len(counts_left)
    q = {k: [] for k in ("valid", "u", "v", "uR", "level", "viewcos", "desc", "obs_pos")}
    for i in range(n):
        c = int(counts_left[i])
        qi = synth.projection_queries(kps_left[i][:c], desc_left[i][:c], w, h, nq, seed0 + i)
        for k in q:
            q[k].append(qi[k])
    return {k: np.ascontiguousarray(np.stack(v)) for k, v in q.items()}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C++ restatement of the reference) -- test/baseline infrastructure, never the product
# ---------------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, threads, w=W, h=H, nfeat=NFEAT):
        import oracle
        self.oracle = oracle
        self.threads = threads
        self.w, self.h, self.nfeat = w, h, nfeat
        try:
            oracle.lib(native=True)          # -O3 -march=native, the reference's own flags (CMakeLists.txt:10-11)
            self.native = True
        except Exception:
            self.native = False
        self.tls = threading.local()
        self.sf = None
        self.with_stereo = False

    def _extractor2(self):
        if not hasattr(self.tls, "ex2"):
            self.tls.ex2 = self.oracle.Extractor(self.nfeat, SCALE, NLEVELS, INI_TH, MIN_TH, native=self.native)
        return self.tls.ex2

    def _extractor(self):
        if not hasattr(self.tls, "ex"):
            self.tls.ex = self.oracle.Extractor(self.nfeat, SCALE, NLEVELS, INI_TH, MIN_TH, native=self.native)
            self.sf = self.tls.ex.scale_factors()
        return self.tls.ex

    def frame(self, left, right, q):
        """one stereo frame: extract L+R (the reference runs them on two threads; here a worker does both,
        all cores being busy with other frames), grid + SearchByProjection on the left frame"""
        ex = self._extractor()
        kl, dl = ex(left)
        if self.with_stereo:
            ex2 = self._extractor2()
            kr, dr = ex2(right)
            _, ur, _ = self.oracle.compute_stereo_matches(ex, ex2, kl, dl, kr, dr, MB, MBF)
            F = self.oracle.Frame(kl, dl, *GRID, u_right=ur, native=self.native)
        else:
            kr, dr = ex(right)
            F = self.oracle.Frame(kl, dl, *GRID, native=self.native)
        n, bi, bd, qk = self.oracle.search_by_projection(F, self.sf, q["valid"], q["u"], q["v"], q["uR"], q["level"], q["viewcos"],
                                                         q["desc"], q["obs_pos"], None, TH, RATIO)
        return len(kl), len(kr), n, bi, bd

    def run(self, imgs, queries, frames):
        """process `frames` (list of frame indices) on self.threads host threads; returns seconds"""
        from concurrent.futures import ThreadPoolExecutor

        def work(i):
            q = {k: v[i] for k, v in queries.items()}
            return self.frame(imgs[2 * i], imgs[2 * i + 1], q)

        t0 = time.perf_counter()
        if self.threads == 1:
            out = [work(i) for i in frames]
        else:
            with ThreadPoolExecutor(self.threads) as pool:
                out = list(pool.map(work, frames))
        return time.perf_counter() - t0, out


def cpu_queries_for(imgs, n_frames, nq, seed0, cpu):
    """queries for the CPU arm built from the oracle's own extraction of the left images"""
    ex = cpu._extractor()
    ks, ds = [], []
    for i in range(n_frames):
        k, d = ex(imgs[2 * i])
        ks.append(k)
        ds.append(d)
    return make_queries(ks, ds, [len(k) for k in ks], nq, seed0)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


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
        # CUDA_VISIBLE_DEVICES may remap: match by UUID through torch
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
    def __init__(self, device, frames_per_step, pools, nq=NQ, w=W, h=H, nfeat=NFEAT, n_ctx=1):
        import torch

        import orb_slam_birdview_b200 as pkg
        self.torch, self.pkg = torch, pkg
        self.device, self.B, self.P, self.nq, self.w, self.h = device, frames_per_step, pools, nq, w, h
        torch.cuda.set_device(device)
        self.ctxs = [pkg.Context(nfeat, SCALE, NLEVELS, INI_TH, MIN_TH, w, h, 2 * frames_per_step, device) for _ in range(n_ctx)]
        self.L = self.ctxs[0]._L
        self.cap = self.ctxs[0].max_keypoints
        self.streams = [torch.cuda.ExternalStream(self.L.orbb200_stream(c._h), device=device) for c in self.ctxs]

    # -- data ------------------------------------------------------------------------------------------
    def setup_data(self, seed0):
        torch = self.torch
        B, P = self.B, self.P
        self.h_imgs, self.h_q, self.d_imgs, self.d_q, self.d_out, self.h_out = [], [], [], [], [], []
        ex_ctx = self.ctxs[0]
        for p in range(P):
            imgs = make_images(B, seed0 + 1000 * p, self.w, self.h)
            # keypoints of the left images (for query generation) from our own, parity-checked extraction
            k = np.empty((2 * B, self.cap), self.pkg.KP_DTYPE)
            d = np.empty((2 * B, self.cap, 32), np.uint8)
            n = np.empty(2 * B, np.int32)
            ptrs = (C.c_void_p * (2 * B))(*[imgs[i].ctypes.data for i in range(2 * B)])
            ex_ctx.check(self.L.orbb200_extract_batch(ex_ctx._h, ptrs, 2 * B, self.w, self.h, self.w, C.c_void_p(k.ctypes.data),
                                                      C.c_void_p(d.ctypes.data), self.cap, C.c_void_p(n.ctypes.data)), "extract_batch")
            q = make_queries(k[0::2], d[0::2], n[0::2], self.nq, seed0 + 1000 * p + 500, self.w, self.h)
            hi = torch.from_numpy(imgs).pin_memory()
            hq = {kk: torch.from_numpy(v).pin_memory() for kk, v in q.items()}
            self.h_imgs.append(hi)
            self.h_q.append(hq)
            self.d_imgs.append(hi.to(f"cuda:{self.device}"))
            self.d_q.append({kk: v.to(f"cuda:{self.device}") for kk, v in hq.items()})
            self.d_out.append(dict(bi=torch.empty((B, self.nq), dtype=torch.int32, device=f"cuda:{self.device}"),
                                   bd=torch.empty((B, self.nq), dtype=torch.int32, device=f"cuda:{self.device}"),
                                   nm=torch.empty(B, dtype=torch.int32, device=f"cuda:{self.device}")))
        # host result buffers, one set per context (e2e)
        for _ in self.ctxs:
            self.h_out.append(dict(
                kps=torch.empty((2 * B, self.cap, 28), dtype=torch.uint8).pin_memory(),
                desc=torch.empty((2 * B, self.cap, 32), dtype=torch.uint8).pin_memory(),
                counts=torch.empty(2 * B, dtype=torch.int32).pin_memory(),
                bi=torch.empty((B, self.nq), dtype=torch.int32).pin_memory(),
                bd=torch.empty((B, self.nq), dtype=torch.int32).pin_memory(),
                nm=torch.empty(B, dtype=torch.int32).pin_memory()))
        torch.cuda.synchronize()

    def _qstruct(self, q):
        s = self.pkg.ProjQueries()
        s.q_valid, s.q_u, s.q_v, s.q_uR = q["valid"].data_ptr(), q["u"].data_ptr(), q["v"].data_ptr(), q["uR"].data_ptr()
        s.q_level, s.q_viewcos, s.q_desc, s.q_obs_pos = q["level"].data_ptr(), q["viewcos"].data_ptr(), q["desc"].data_ptr(), q["obs_pos"].data_ptr()
        return s

    # -- steps -----------------------------------------------------------------------------------------
    def step_device(self, p, ctx_i=0):
        ctx = self.ctxs[ctx_i]
        qs = self._qstruct(self.d_q[p])
        o = self.d_out[p]
        ctx.check(self.L.orbb200_stereo_step_device(ctx._h, self.d_imgs[p].data_ptr(), self.w * self.h, self.B, self.w, self.h, self.w,
                                                    C.byref(qs), self.nq, TH, RATIO, *GRID, o["bi"].data_ptr(), o["bd"].data_ptr(),
                                                    o["nm"].data_ptr()), "stereo_step_device")

    def step_host(self, p, ctx_i=0):
        ctx = self.ctxs[ctx_i]
        qs = self._qstruct(self.h_q[p])
        o = self.h_out[ctx_i]
        ctx.check(self.L.orbb200_stereo_step_host(ctx._h, self.h_imgs[p].data_ptr(), self.B, self.w, self.h, self.w, C.byref(qs), self.nq,
                                                  TH, RATIO, *GRID, o["kps"].data_ptr(), o["desc"].data_ptr(), self.cap,
                                                  o["counts"].data_ptr(), o["bi"].data_ptr(), o["bd"].data_ptr(), o["nm"].data_ptr()),
                  "stereo_step_host")

    def h2d_bytes(self):
        q = self.h_q[0]
        return int(self.h_imgs[0].numel() + sum(v.numel() * v.element_size() for v in q.values()))

    def d2h_bytes(self):
        o = self.h_out[0]
        return int(sum(v.numel() * v.element_size() for v in o.values()))

    def launches(self):
        return sum(c.launches for c in self.ctxs)


def parity_check(n_frames=3, nq=600, w=W, h=H, nfeatures=NFEAT, device=0):
    """The batched device step (what the bench times) against the oracle, frame by frame."""
    import oracle
    arm = GpuArm(device, n_frames, 1, nq=nq, w=w, h=h, nfeat=nfeatures)
    arm.setup_data(4242)
    arm.step_device(0)
    arm.step_host(0)
    arm.ctxs[0].sync()
    o = arm.d_out[0]
    bi, bd, nm = o["bi"].cpu().numpy(), o["bd"].cpu().numpy(), o["nm"].cpu().numpy()
    ho = arm.h_out[0]
    imgs = arm.h_imgs[0].numpy()
    q = {k: v.numpy() for k, v in arm.h_q[0].items()}
    orc = oracle.Extractor(nfeatures, SCALE, NLEVELS, INI_TH, MIN_TH)
    res = {"ok": True, "frames": []}
    for i in range(n_frames):
        kl, dl = orc(imgs[2 * i])
        kr, dr = orc(imgs[2 * i + 1])
        F = oracle.Frame(kl, dl, np.float32(0), np.float32(0), np.float32(64.0 / w), np.float32(48.0 / h))
        n0, bi0, bd0, _ = oracle.search_by_projection(F, orc.scale_factors(), q["valid"][i], q["u"][i], q["v"][i], q["uR"][i], q["level"][i],
                                                      q["viewcos"][i], q["desc"][i], q["obs_pos"][i], None, TH, RATIO)
        cnt = ho["counts"].numpy()
        kps_l = ho["kps"].numpy()[2 * i].view(arm.pkg.KP_DTYPE).reshape(-1)[:cnt[2 * i]]
        kps_r = ho["kps"].numpy()[2 * i + 1].view(arm.pkg.KP_DTYPE).reshape(-1)[:cnt[2 * i + 1]]
        ok = (int(nm[i]) == n0 and np.array_equal(bi[i], bi0) and np.array_equal(bd[i][bi0 >= 0], bd0[bi0 >= 0])
              and kps_l.tobytes() == kl.tobytes() and kps_r.tobytes() == kr.tobytes()
              and np.array_equal(ho["desc"].numpy()[2 * i][:len(kl)], dl) and np.array_equal(ho["desc"].numpy()[2 * i + 1][:len(kr)], dr)
              and np.array_equal(ho["bi"].numpy()[i], bi0) and int(ho["nm"].numpy()[i]) == n0)
        res["frames"].append({"nmatches": int(nm[i]), "oracle": int(n0), "ok": bool(ok)})
        res["ok"] = res["ok"] and bool(ok)
    return res


def dist_setup(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local)
        dist_.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        dist = dist_
    return rank, world, local, dist


_FULL_AFFINITY = None      # the process's CPU set before bind_near_gpu narrowed it (the CPU baseline gets all of it back)


def bind_near_gpu(device, enable=True):
    """Pin this process (and the threads / pinned host allocations it makes afterwards) to the CPUs NVML reports as
    closest to the GPU, so that each rank's staging buffers live on the NUMA node its GPU's PCIe root hangs off.
    Returns a small description for the JSON line."""
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


def run_ours(args):
    import torch
    rank, world, local, dist = dist_setup(args)
    host_bind = bind_near_gpu(local, not args.no_numa_bind)
    B, P = args.frames_per_step, args.pools
    NE = max(2, args.e2e_contexts)
    arm = GpuArm(local, B, P, n_ctx=NE)
    arm.setup_data(100000 * rank + 2000)
    if args.with_stereo:
        for c in arm.ctxs:
            c.check(arm.L.orbb200_step_enable_stereo(c._h, 1, MB, MBF), "step_enable_stereo")
    K, Wm = args.steps, max(args.warmup, 3)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (value): CUDA events on the contexts' streams ----
    # Steps alternate over --device-contexts contexts (streams), default 2: the tail of one step (matching: a few
    # hundred latency-bound CTAs) then overlaps the head of the next.  Measured 2.53 ms per step against 2.60 on one.
    nctx = max(1, min(args.device_contexts, NE))
    # warm-up visits every (input batch, context) pair the timed region uses: the first use of a pair builds its step plan
    # (cudaMalloc + graph capture), which must not happen inside the timed region
    Wm = max(Wm, P * nctx)
    for s in range(Wm):
        arm.step_device(s % P, s % nctx)
    for c in arm.ctxs:
        c.sync()
    for c in arm.ctxs[:nctx]:
        arm.L.orbb200_stage_timing(c._h, 1 if args.stage_timing_in_region else 0)
        arm.L.orbb200_stage_times(c._h, None, None, 1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = arm.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    join = torch.cuda.Event()
    sampler.mark()
    e0.record(arm.streams[0])
    for i in range(1, nctx):
        arm.streams[i].wait_event(e0)
    for s in range(K):
        arm.step_device(s % P, s % nctx)
    for i in range(1, nctx):
        join = torch.cuda.Event()
        join.record(arm.streams[i])
        arm.streams[0].wait_event(join)
    e1.record(arm.streams[0])
    barrier()
    sampler.mark()
    ms = e0.elapsed_time(e1)
    launches = arm.launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    if not args.stage_timing_in_region:
        # per-stage durations from a second pass of the same K steps with events between the stages
        # on ONE context, so that a stage's duration is that of its kernels running alone
        nctx_t = 1
        for c in arm.ctxs[:nctx_t]:
            arm.L.orbb200_stage_timing(c._h, 1)
            arm.L.orbb200_stage_times(c._h, None, None, 1)
        for s in range(K):
            arm.step_device(s % P, s % nctx_t)
        for c in arm.ctxs[:nctx_t]:
            c.sync()
    else:
        nctx_t = nctx
    st_ms = np.zeros(9, np.float32)
    st_n = np.zeros(9, np.int32)
    for c in arm.ctxs[:nctx_t]:
        a_ms = np.zeros(9, np.float32)
        a_n = np.zeros(9, np.int32)
        arm.L.orbb200_stage_times(c._h, C.c_void_p(a_ms.ctypes.data), C.c_void_p(a_n.ctypes.data), 1)
        arm.L.orbb200_stage_timing(c._h, 0)
        st_ms += a_ms
        st_n += a_n

    # ---- end to end through the host-buffer C-ABI call: NE contexts in flight ----
    for s in range(max(Wm, 2 * NE, P * NE) if not args.no_e2e else 0):      # every (input batch, context) pair once
        arm.step_host(s % P, s % NE)
    for c in arm.ctxs:
        c.sync()
    barrier()
    t0 = time.perf_counter()
    for s in range(K if not args.no_e2e else 0):
        ci = s % NE
        if s >= NE:
            arm.ctxs[ci].sync()          # results of step s-NE are on the host: consume before reuse
            _ = int(arm.h_out[ci]["nm"][0])
        arm.step_host(s % P, ci)
    for c in arm.ctxs:
        c.sync()
    _ = int(arm.h_out[0]["nm"][0])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # host->device copy rate with every rank copying at the same time (pinned, one stream each): what the host side
    # (PCIe switches, NUMA placement of the staging buffers) gives N GPUs together; the floor under the host-buffer leg
    ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    arm.d_imgs[0].copy_(arm.h_imgs[0], non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    ce0.record()
    for _ in range(4):
        arm.d_imgs[0].copy_(arm.h_imgs[0], non_blocking=True)
    ce1.record()
    torch.cuda.synchronize()
    h2d_conc = 4 * arm.h_imgs[0].numel() / (ce0.elapsed_time(ce1) * 1e-3) / 1e9
    h2d_conc_min, h2d_conc_sum = h2d_conc, h2d_conc
    barrier()

    if dist is not None:
        t = torch.tensor([ms, e2e_s * 1e3, -h2d_conc], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, h2d_conc_min = float(t[0]), float(t[1]), -float(t[2])
        t2 = torch.tensor([h2d_conc], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.SUM)
        h2d_conc_sum = float(t2[0])
    else:
        e2e_ms = e2e_s * 1e3
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # host->device copy rate of this box (pinned, one stream): the floor under the host-buffer leg
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    arm.d_imgs[0].copy_(arm.h_imgs[0], non_blocking=True)
    pe0.record()
    for _ in range(4):
        arm.d_imgs[0].copy_(arm.h_imgs[0], non_blocking=True)
    pe1.record()
    torch.cuda.synchronize()
    h2d_gbs = 4 * arm.h_imgs[0].numel() / (pe0.elapsed_time(pe1) * 1e-3) / 1e9

    frames_total = world * B * K
    value = frames_total / (ms * 1e-3)
    e2e_value = frames_total / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel ----
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
    alg = {"pyramid": 2 * sum_p - p0 - level_pixels(W, H, last=True), "fast": sum_p, "blur": 2 * sum_p,
           "import": 2 * p0, "octree": None, "describe": None, "grid": None, "match": None}
    name = STAGES[dom]
    bytes_per_launch = alg.get(name)
    traffic = None
    ncu_pipes = {}
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

    # ---- CPU baseline on this box's host cores: bounded sample of the same workload ----
    cpu = cpu_baseline(arm, args) if not args.no_cpu_baseline else None

    out = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "C2: KITTI-shape stereo 1241x376 L+R, 2000 feats/img, 8 levels, scale 1.2, FAST 20/7, extract + "
                               "SearchByProjection vs 3000 local map points (th=1, ratio 0.8)",
                   "frames_per_step_per_gpu": B, "queries_per_frame": NQ, "sharding": f"frames x{world} (no collective on the path)",
                   "l2": f"{P} rotating input batches of {2 * B} images + {2 * B}-image pyramid/blur pools: working set "
                         f"{working_set_mb(B, P):.0f} MB per GPU > 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": arm.h2d_bytes(), "d2h_bytes_per_step": arm.d2h_bytes(),
                "ms_per_step": e2e_ms / K, "contexts_in_flight": NE, "h2d_copy_gbs_measured": h2d_gbs,
                "h2d_copy_gbs_all_ranks_at_once": {"min_rank": h2d_conc_min, "sum": h2d_conc_sum}, "host_binding": host_bind},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "stage_ms_per_step": {STAGES[i]: float(st_ms[i] / K) for i in range(9)},
        "stage_timing": ("CUDA events between the stages inside the timed region" if args.stage_timing_in_region else
                         "CUDA events between the stages in a second pass of the same K steps on one context (the timed region replays the captured graphs, blur forked beside FAST)"),
        "device_contexts": nctx,
        "cpu_baseline": cpu,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def level_pixels(w, h, last=False):
    s = np.float32(1.0)
    tot, p0, pl = 0, w * h, 0
    for l in range(NLEVELS):
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


def cpu_baseline(arm, args, target_s=15.0):
    """Oracle port on the host cores; bounded sample (about target_s seconds of wall time)."""
    try:
        if _FULL_AFFINITY:
            os.sched_setaffinity(0, _FULL_AFFINITY)      # the baseline may use every host core, not only the GPU's neighbours
        cores = host_cores()
        cpu = CpuArm(cores)
        cpu.with_stereo = bool(args.with_stereo)
        imgs = arm.h_imgs[0].numpy()
        q = {k: v.numpy() for k, v in arm.h_q[0].items()}
        navail = imgs.shape[0] // 2
        t1, _ = CpuArm(1).run(imgs, q, [0])                      # one frame, one thread: sizes the sample
        t1b, _ = CpuArm(1).run(imgs, q, [0])
        t1 = min(t1, t1b)
        n = int(min(max(cores * 2, target_s * cores / max(t1, 1e-3)), 400 * cores))
        frames = [i % navail for i in range(n)]
        cpu.run(imgs, q, frames[:cores])                          # warm-up
        dt, _ = cpu.run(imgs, q, frames)
        return {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port",
                "sample": f"{n} stereo frames of the same workload on {cores} host threads ({dt:.1f} s); 1 thread: {1.0 / t1:.2f} frames/s",
                "single_thread_value": 1.0 / t1, "native_build": cpu.native,
                "cv2_crosscheck": cv2_primitive_times(imgs[0], 1e3 * t1 / 2)}
    except Exception as e:   # the baseline is reported, never required for the GPU number
        return {"value": None, "unit": "frames/s", "cores": 0, "kind": "port", "sample": f"failed: {e!r}"}


def cv2_primitive_times(img, oracle_ms_per_image):
    """SURVEY.md 8(d) cross-check of the scalar oracle port against OpenCV's own SIMD builds of the three stencil
    stages (cv2.resize, per-cell cv2.FAST with the 20/7 fallback, cv2.GaussianBlur), one thread, one image.  The
    per-cell FAST calls are made from Python, so the cost of the same number of calls on a 7x7 cell (no interior pixel
    to test) is measured and subtracted.  Octree, orientation and descriptors have no cv2 counterpart and are not in
    this sum, so it is a LOWER bound of an OpenCV-based CPU extraction; reported next to the oracle's time per image."""
    try:
        import cv2
        import numpy as np
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
    stencil = r[0] + max(r[1] - r[3], 0.0) + r[2]
    return {"available": True, "version": cv2.__version__, "threads": 1, "resize_ms": round(r[0], 3), "fast_cells_ms": round(r[1], 3),
            "fast_call_overhead_ms": round(r[3], 3), "blur_ms": round(r[2], 3), "stencil_stages_ms_per_image": round(stencil, 3),
            "oracle_ms_per_image_all_stages": round(oracle_ms_per_image, 3),
            "note": "cv2 covers pyramid + FAST + blur only (no octree / orientation / descriptors / matching)"}


def run_reference(args):
    """--impl reference: the reference algorithm (C++ oracle port; the reference itself needs OpenCV/Eigen/
    Pangolin SDKs that are not installed) on all host threads, same workload/metric."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = host_cores()
    cpu = CpuArm(cores)
    F = max(4 * cores, 16) if args.ref_frames_per_step <= 0 else args.ref_frames_per_step   # 4 frames per thread per step: keeps every core busy
    npool = min(F, 32)
    imgs = make_images(npool, 2000)
    q = cpu_queries_for(imgs, npool, NQ, 2500, cpu)
    frames = [i % npool for i in range(F)]
    K, Wm = args.steps, max(args.warmup, 1)
    for _ in range(Wm):
        cpu.run(imgs, q, frames)
    t0 = time.perf_counter()
    for _ in range(K):
        cpu.run(imgs, q, frames)
    dt = time.perf_counter() - t0
    value = F * K / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C2: KITTI-shape stereo 1241x376 L+R, 2000 feats/img, 8 levels, scale 1.2, FAST 20/7, extract + "
                               "SearchByProjection vs 3000 local map points (th=1, ratio 0.8)",
                   "frames_per_step": F, "queries_per_frame": NQ},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{F} stereo frames per step on {cores} host threads, native_build={cpu.native}"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=128)
    ap.add_argument("--pools", type=int, default=3)
    ap.add_argument("--ref-frames-per-step", type=int, default=0)
    ap.add_argument("--with-stereo", action="store_true", help="also run ComputeStereoMatches in the step (side measurement; not the C2 headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the host-core baseline (profiling runs)")
    ap.add_argument("--device-contexts", type=int, default=2, help="contexts (streams) the device-resident leg alternates its steps over")
    ap.add_argument("--e2e-contexts", type=int, default=3, help="contexts (streams) the host-buffer leg keeps in flight")
    ap.add_argument("--stage-timing-in-region", action="store_true",
                    help="record the per-stage events inside the timed region (plain launches, no graph replay, blur not forked: ~2.5 %% slower); "
                         "default: the timed region replays the captured graph and a second pass of the same K steps times the stages")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the rank to the CPUs nearest its GPU")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
