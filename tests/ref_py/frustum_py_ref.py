"""Independent transcription of Frame::isInFrustum + MapPoint::PredictScale (reference src/Frame.cc:436-492,
src/MapPoint.cc:402-417) on top of the real OpenCV primitives the reference calls: cv2.gemm for mRcw*P+mtcw and
cv2.norm for |P-Ow| (cv2 4.13, the only runnable OpenCV here).  Mat::dot has no Python binding; it accumulates
in double, which float64 numpy reproduces.  Test infrastructure: pins the oracle's restatement of the cv::Mat
arithmetic.  Pure-Python loop: small cases only."""
import math

import cv2
import numpy as np

f32 = np.float32


def is_in_frustum(pos, normal, max_distance, min_distance, pose, viewing_cos_limit=0.5, candidate=None):
    R = np.array(pose["Rcw"], f32).reshape(3, 3)
    t = np.array(pose["tcw"], f32).reshape(3, 1)
    Ow = np.array(pose["Ow"], f32).reshape(3, 1)
    fx, fy, cx, cy, mbf = (f32(pose[k]) for k in ("fx", "fy", "cx", "cy", "mbf"))
    n = len(pos)
    out = dict(in_view=np.zeros(n, np.uint8), u=np.zeros(n, f32), v=np.zeros(n, f32), uR=np.zeros(n, f32),
               level=np.zeros(n, np.int32), viewcos=np.zeros(n, f32))
    for i in range(n):
        if candidate is not None and not candidate[i]:
            continue
        P = np.array(pos[i], f32).reshape(3, 1)
        Pc = cv2.gemm(R, P, 1.0, t, 1.0)
        PcX, PcY, PcZ = f32(Pc[0, 0]), f32(Pc[1, 0]), f32(Pc[2, 0])
        if PcZ < 0:
            continue
        with np.errstate(all="ignore"):
            invz = f32(1.0) / PcZ
            u = f32(f32(fx * PcX) * invz) + cx
            v = f32(f32(fy * PcY) * invz) + cy
        if u < pose["min_x"] or u > pose["max_x"] or v < pose["min_y"] or v > pose["max_y"]:
            continue
        maxd = f32(1.2) * f32(max_distance[i])
        mind = f32(0.8) * f32(min_distance[i])
        PO = (P - Ow).astype(f32)
        dist = f32(cv2.norm(PO))
        if dist < mind or dist > maxd:
            continue
        Pn = np.array(normal[i], f32).reshape(3, 1)
        dot = float(PO[0, 0]) * float(Pn[0, 0]) + float(PO[1, 0]) * float(Pn[1, 0]) + float(PO[2, 0]) * float(Pn[2, 0])
        viewcos = f32(dot / float(dist))
        if viewcos < f32(viewing_cos_limit):
            continue
        ratio = f32(max_distance[i]) / dist
        lg = f32(math.log(float(ratio)))                     # logf, correctly rounded
        c = math.ceil(float(lg / f32(pose["log_scale_factor"])))
        lvl = min(max(c, 0), pose["n_levels"] - 1)
        out["in_view"][i] = 1
        out["u"][i], out["v"][i], out["uR"][i] = u, v, u - f32(mbf * invz)
        out["level"][i], out["viewcos"][i] = lvl, viewcos
    return out
