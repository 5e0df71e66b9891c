"""Independent Python restatement of ORB_SLAM2::ORBextractor built on cv2 primitives.

Purpose: pin the C++ oracle (oracle/orb_oracle.cpp).  Every OpenCV call the reference makes is made here
through the real OpenCV (cv2 4.13: resize, FAST on the per-cell sub-image, GaussianBlur, fastAtan2);
the reference's own code (src/ORBextractor.cc) is transcribed with Python lists.  It is only run in the
build container by tests/golden/make_golden.py (cv2 may be absent elsewhere); the vectors it produced
are committed under tests/golden/.

Tie rule for the size sort at src/ORBextractor.cc:684 (pair<int,ExtractorNode*>): most recently created
node first among equal sizes (see oracle/orb_oracle.h).
"""
import ctypes
import ctypes.util
import math

import cv2
import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library("m"))
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]

F32 = np.float32
PATCH_SIZE, HALF_PATCH_SIZE, EDGE_THRESHOLD = 31, 15, 19


def cv_round(v):
    """cvRound: round half to even"""
    return int(np.rint(v))


def load_pattern():
    import os
    import re
    txt = open(os.path.join(os.path.dirname(__file__), "..", "..", "include", "orbb200_pattern.inc")).read()
    xs = re.search(r"PATTERN_X_INIT(.*?)\n\n", txt, re.S).group(1)
    ys = txt[txt.index("PATTERN_Y_INIT") + len("PATTERN_Y_INIT"):]
    gx = [int(v) for v in re.findall(r"-?\d+", xs)]
    gy = [int(v) for v in re.findall(r"-?\d+", ys)]
    assert len(gx) == 512 and len(gy) == 512
    return gx, gy


class Node:
    __slots__ = ("keys", "UL", "UR", "BL", "BR", "no_more", "seq", "alive")

    def __init__(self):
        self.keys = []
        self.no_more = False
        self.seq = 0
        self.alive = True


def divide_node(n):
    half_x = int(math.ceil(float(F32(n.UR[0] - n.UL[0]) / F32(2))))
    half_y = int(math.ceil(float(F32(n.BR[1] - n.UL[1]) / F32(2))))
    n1, n2, n3, n4 = Node(), Node(), Node(), Node()
    n1.UL = n.UL
    n1.UR = (n.UL[0] + half_x, n.UL[1])
    n1.BL = (n.UL[0], n.UL[1] + half_y)
    n1.BR = (n.UL[0] + half_x, n.UL[1] + half_y)
    n2.UL = n1.UR
    n2.UR = n.UR
    n2.BL = n1.BR
    n2.BR = (n.UR[0], n.UL[1] + half_y)
    n3.UL = n1.BL
    n3.UR = n1.BR
    n3.BL = n.BL
    n3.BR = (n1.BR[0], n.BL[1])
    n4.UL = n3.UR
    n4.UR = n2.BR
    n4.BL = n3.BR
    n4.BR = n.BR
    for kp in n.keys:
        if kp[0] < n1.UR[0]:
            (n1 if kp[1] < n1.BR[1] else n3).keys.append(kp)
        elif kp[1] < n1.BR[1]:
            n2.keys.append(kp)
        else:
            n4.keys.append(kp)
    for c in (n1, n2, n3, n4):
        if len(c.keys) == 1:
            c.no_more = True
    return n1, n2, n3, n4


def distribute_octree(keys, min_x, max_x, min_y, max_y, N):
    """keys: list of (x, y, response) in region coordinates.  Python list stands in for std::list:
    index 0 is the list front."""
    n_ini = int(np.round(F32(max_x - min_x) / F32(max_y - min_y)))
    h_x = F32(max_x - min_x) / F32(n_ini)
    seq = 0
    nodes = []
    for i in range(n_ini):
        ni = Node()
        ni.UL = (int(h_x * F32(i)), 0)
        ni.UR = (int(h_x * F32(i + 1)), 0)
        ni.BL = (ni.UL[0], max_y - min_y)
        ni.BR = (ni.UR[0], max_y - min_y)
        ni.seq = seq
        seq += 1
        nodes.append(ni)
    ini = list(nodes)
    for kp in keys:
        ini[int(F32(kp[0]) / h_x)].keys.append(kp)
    kept = []
    for n in nodes:
        if len(n.keys) == 1:
            n.no_more = True
            kept.append(n)
        elif len(n.keys) > 0:
            kept.append(n)
    nodes = kept

    finish = False
    while not finish:
        prev_size = len(nodes)
        n_to_expand = 0
        size_nodes = []
        front = []          # pushed to the list front, most recent first
        rest = []
        for n in nodes:
            if n.no_more:
                rest.append(n)
                continue
            for c in divide_node(n):
                if len(c.keys) > 0:
                    c.seq = seq
                    seq += 1
                    front.insert(0, c)
                    if len(c.keys) > 1:
                        n_to_expand += 1
                        size_nodes.append(c)
        nodes = front + rest
        if len(nodes) >= N or len(nodes) == prev_size:
            finish = True
        elif len(nodes) + n_to_expand * 3 > N:
            while not finish:
                prev_size = len(nodes)
                prev_nodes = size_nodes
                size_nodes = []
                prev_nodes = sorted(prev_nodes, key=lambda c: (len(c.keys), c.seq))
                for j in range(len(prev_nodes) - 1, -1, -1):
                    n = prev_nodes[j]
                    for c in divide_node(n):
                        if len(c.keys) > 0:
                            c.seq = seq
                            seq += 1
                            nodes.insert(0, c)
                            if len(c.keys) > 1:
                                size_nodes.append(c)
                    nodes.remove(n)
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prev_size:
                    finish = True
    out = []
    for n in nodes:
        best = n.keys[0]
        for kp in n.keys[1:]:
            if kp[2] > best[2]:
                best = kp
        out.append(best)
    return out


class OrbCv2Ref:
    def __init__(self, nfeatures, scale_factor, nlevels, ini_th, min_th):
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sf_d = float(F32(scale_factor))          # member is double, initialised from the float argument
        self.scale = [F32(1.0)]
        for i in range(1, nlevels):
            self.scale.append(F32(float(self.scale[i - 1]) * sf_d))
        self.inv_scale = [F32(1.0) / s for s in self.scale]
        factor = F32(1.0 / sf_d)
        nd = F32(nfeatures) * (F32(1) - factor) / (F32(1) - F32(math.pow(float(factor), float(nlevels))))
        self.n_per_level = []
        tot = 0
        for _ in range(nlevels - 1):
            self.n_per_level.append(cv_round(nd))
            tot += self.n_per_level[-1]
            nd = F32(nd * factor)
        self.n_per_level.append(max(nfeatures - tot, 0))
        umax = [0] * (HALF_PATCH_SIZE + 1)
        vmax = int(math.floor(float(F32(HALF_PATCH_SIZE) * F32(math.sqrt(2.0)) / F32(2) + F32(1))))
        vmin = int(math.ceil(float(F32(HALF_PATCH_SIZE) * F32(math.sqrt(2.0)) / F32(2))))
        hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE
        for v in range(vmax + 1):
            umax[v] = cv_round(math.sqrt(hp2 - v * v))
        v0 = 0
        for v in range(HALF_PATCH_SIZE, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
        self.umax = umax
        self.pat_x, self.pat_y = load_pattern()

    def pyramid(self, img):
        pyr = []
        for lvl in range(self.nlevels):
            s = self.inv_scale[lvl]
            sz = (cv_round(F32(img.shape[1]) * s), cv_round(F32(img.shape[0]) * s))
            if lvl == 0:
                pyr.append(img.copy())
            else:
                pyr.append(cv2.resize(pyr[lvl - 1], sz, interpolation=cv2.INTER_LINEAR))
        return pyr

    def ic_angle(self, im, x, y):
        cx, cy = cv_round(x), cv_round(y)
        m01 = m10 = 0
        for u in range(-HALF_PATCH_SIZE, HALF_PATCH_SIZE + 1):
            m10 += u * int(im[cy, cx + u])
        for v in range(1, HALF_PATCH_SIZE + 1):
            v_sum = 0
            d = self.umax[v]
            for u in range(-d, d + 1):
                vp, vm = int(im[cy + v, cx + u]), int(im[cy - v, cx + u])
                v_sum += vp - vm
                m10 += u * (vp + vm)
            m01 += v * v_sum
        return cv2.fastAtan2(float(m01), float(m10))

    def descriptor(self, im, x, y, angle_deg):
        factor_pi = F32(math.pi / float(F32(180.0)))
        ang = F32(F32(angle_deg) * factor_pi)
        a, b = F32(_libm.cosf(float(ang))), F32(_libm.sinf(float(ang)))
        cx, cy = cv_round(x), cv_round(y)
        out = np.zeros(32, np.uint8)

        def val(i):
            px, py = F32(self.pat_x[i]), F32(self.pat_y[i])
            yy = cv_round(F32(F32(px * b) + F32(py * a)))
            xx = cv_round(F32(F32(px * a) - F32(py * b)))
            return int(im[cy + yy, cx + xx])

        for i in range(32):
            v = 0
            for k in range(8):
                if val(16 * i + 2 * k) < val(16 * i + 2 * k + 1):
                    v |= 1 << k
            out[i] = v
        return out

    def __call__(self, img):
        pyr = self.pyramid(img)
        fast_ini = cv2.FastFeatureDetector_create(self.ini_th, True)
        fast_min = cv2.FastFeatureDetector_create(self.min_th, True)
        all_kps = []
        W = F32(30)
        for lvl in range(self.nlevels):
            im = pyr[lvl]
            min_bx = min_by = EDGE_THRESHOLD - 3
            max_bx = im.shape[1] - EDGE_THRESHOLD + 3
            max_by = im.shape[0] - EDGE_THRESHOLD + 3
            width, height = F32(max_bx - min_bx), F32(max_by - min_by)
            n_cols, n_rows = int(width / W), int(height / W)
            w_cell = int(math.ceil(float(width / F32(n_cols))))
            h_cell = int(math.ceil(float(height / F32(n_rows))))
            cand = []
            for i in range(n_rows):
                ini_y = min_by + i * h_cell
                max_y = ini_y + h_cell + 6
                if ini_y >= max_by - 3:
                    continue
                max_y = min(max_y, max_by)
                for j in range(n_cols):
                    ini_x = min_bx + j * w_cell
                    max_x = ini_x + w_cell + 6
                    if ini_x >= max_bx - 6:
                        continue
                    max_x = min(max_x, max_bx)
                    cell = np.ascontiguousarray(im[ini_y:max_y, ini_x:max_x])
                    kps = fast_ini.detect(cell)
                    if len(kps) == 0:
                        kps = fast_min.detect(cell)
                    for kp in kps:
                        cand.append((float(F32(kp.pt[0]) + F32(j * w_cell)), float(F32(kp.pt[1]) + F32(i * h_cell)),
                                     float(kp.response)))
            keys = distribute_octree(cand, min_bx, max_bx, min_by, max_by, self.n_per_level[lvl])
            size = float(int(F32(PATCH_SIZE) * self.scale[lvl]))
            lv = []
            for (x, y, r) in keys:
                lv.append([x + min_bx, y + min_by, size, -1.0, r, lvl, -1])
            all_kps.append(lv)
        for lvl in range(self.nlevels):
            for k in all_kps[lvl]:
                k[3] = self.ic_angle(pyr[lvl], k[0], k[1])
        out_k, out_d = [], []
        for lvl in range(self.nlevels):
            if not all_kps[lvl]:
                continue
            blur = cv2.GaussianBlur(pyr[lvl], (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
            s = self.scale[lvl]
            for k in all_kps[lvl]:
                out_d.append(self.descriptor(blur, k[0], k[1], k[3]))
                x, y = F32(k[0]), F32(k[1])
                if lvl != 0:
                    x, y = F32(x * s), F32(y * s)
                out_k.append((x, y, k[2], k[3], k[4], k[5], k[6]))
        kp_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                             ("octave", "<i4"), ("class_id", "<i4")])
        kps = np.array(out_k, dtype=kp_dtype) if out_k else np.empty(0, kp_dtype)
        desc = np.stack(out_d) if out_d else np.empty((0, 32), np.uint8)
        return kps, desc
