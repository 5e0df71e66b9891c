#!/usr/bin/env python3
"""Side measurements for the other BASELINE.json configs (C1, C3, C4, C5); bench.py stays the C2 headline.
Writes one JSON object per config to stdout (and gpurun_out/configs.json when that directory exists)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import orb_slam_birdview_b200 as pkg  # noqa: E402
from orb_slam_birdview_b200 import synth  # noqa: E402


def ev_time(stream, fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def extract_config(name, w, h, nfeat, ini, mn, batch, reps=10):
    L = pkg.load_library()
    ctx = pkg.Context(nfeat, 1.2, 8, ini, mn, w, h, batch)
    stream = torch.cuda.ExternalStream(L.orbb200_stream(ctx._h))
    imgs = np.stack([synth.synth_frame(h, w, 7000 + i) for i in range(min(batch, 16))])
    imgs = np.concatenate([imgs] * ((batch + len(imgs) - 1) // len(imgs)))[:batch]
    d = torch.from_numpy(imgs).cuda()
    ms = ev_time(stream, lambda: ctx.check(L.orbb200_extract_device(ctx._h, d.data_ptr(), w * h, batch, w, h, w)), reps)
    # single-image latency through the host API
    ex = pkg.ORBextractor(nfeat, 1.2, 8, ini, mn, max_size=(w, h))
    ex(imgs[0])
    t0 = time.perf_counter()
    for _ in range(20):
        k, _d = ex(imgs[0])
    lat = (time.perf_counter() - t0) / 20
    return {"config": name, "shape": [w, h], "nfeatures": nfeat, "batch": batch, "frames_per_s_device": batch / (ms * 1e-3),
            "ms_per_batch": ms, "single_image_host_api_ms": lat * 1e3, "keypoints": int(len(k))}


def knn2_config(nq, nm, reps=10):
    L = pkg.load_library()
    ctx = pkg.Context(1000, 1.2, 8, 20, 7, 64, 64)
    stream = torch.cuda.ExternalStream(L.orbb200_stream(ctx._h))
    q = torch.from_numpy(synth.synth_descriptors(nq, 1)).cuda()
    m = torch.from_numpy(synth.synth_descriptors(nm, 2)).cuda()
    bi = torch.empty(nq, dtype=torch.int32, device="cuda")
    bd = torch.empty_like(bi)
    sd = torch.empty_like(bi)
    ms = ev_time(stream, lambda: ctx.check(L.orbb200_hamming_knn2_device(ctx._h, q.data_ptr(), nq, m.data_ptr(), nm, bi.data_ptr(),
                                                                         bd.data_ptr(), sd.data_ptr())), reps)
    peak = ctx.popc_peak_gops()
    pairs = nq * nm
    return {"config": f"C4 knn2 {nq}x{nm}", "ms": ms, "gpairs_per_s": pairs / (ms * 1e-3) * 1e-9, "popc_gops": 8 * pairs / (ms * 1e-3) * 1e-9,
            "popc_peak_gops_measured": peak, "popc_frac": 8 * pairs / (ms * 1e-3) * 1e-9 / peak}


def c3_config(batch=32, reps=5):
    """C3: front 1241x376 (ORBextractor, 2000 features) + birdview 400x400 through the reference's real birdview front-end
    (cv::ORB(2000) detect(mask) + cornerSubPix + compute, src/Frame.cc:328-342) + BirdviewMatch of consecutive birdview
    frames (window 15, src/Tracking.cc:326).  Host buffers in and out (wall clock)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    w, h = 1241, 376
    ctx = pkg.Context(2000, 1.2, 8, 20, 7, w, h, batch)
    L = ctx._L
    front = [synth.synth_frame(h, w, 8000 + i) for i in range(batch)]
    base, mask = cases.birdview_case(400, 8100)
    birds = [synth.shift_frame(base, 2 * i % 5, -(i % 4)) for i in range(batch)]
    masks = [mask] * batch
    B = pkg.BirdviewORB(ctx, 2000)
    M = pkg.ORBmatcher(ctx, 0.99, True)
    cap = ctx.max_keypoints
    k = np.empty((batch, cap), pkg.KP_DTYPE)
    d = np.empty((batch, cap, 32), np.uint8)
    n = np.empty(batch, np.int32)
    ptrs = (C.c_void_p * batch)(*[f.ctypes.data for f in front])
    grid = (0.0, 0.0, 64.0 / 400, 48.0 / 400)

    def step():
        ctx.check(L.orbb200_extract_batch(ctx._h, ptrs, batch, w, h, w, C.c_void_p(k.ctypes.data), C.c_void_p(d.ctypes.data), cap,
                                          C.c_void_p(n.ctypes.data)), "extract_batch")
        bk, bd = B.extract_batch(birds, masks)
        nm = 0
        for i in range(1, batch):
            F2 = pkg.Frame(ctx, bk[i], bd[i], *grid)
            nm += M.BirdviewMatch(bk[i - 1], bd[i - 1], F2, 15)[0]
            F2.close()
        return nm, sum(len(x) for x in bk)

    step()
    t0 = time.perf_counter()
    for _ in range(reps):
        nm, nb = step()
    dt = (time.perf_counter() - t0) / reps
    return {"config": "C3 front 1241x376 + birdview 400x400 (cv::ORB + cornerSubPix front-end) + BirdviewMatch", "batch": batch,
            "frames_per_s_host_api": batch / dt, "ms_per_frame": dt / batch * 1e3, "bird_keypoints_per_frame": nb / batch,
            "bird_matches_per_pair": nm / (batch - 1)}


def main():
    out = [extract_config("C1 EuRoC mono", 752, 480, 1000, 20, 7, 64),
           extract_config("C3 birdview image 400x400 (ORBextractor kernels)", 400, 400, 2000, 15, 5, 64),
           extract_config("C5 1080p", 1920, 1080, 4000, 20, 7, 32)]
    out.append(c3_config())
    for nm in (2000, 20000, 200000):
        out.append(knn2_config(2000, nm))
    for o in out:
        print(json.dumps(o))
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
