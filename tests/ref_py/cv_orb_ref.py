"""Transcription of cv::ORB (OpenCV 4.x modules/features2d/src/orb.cpp, HARRIS_SCORE, WTA_K 2, patch 31) as the reference
uses it for the birdview image (src/Frame.cc:328-342: ORB::create(2000)->detect(img, kps, mask); cornerSubPix;
->compute), with the OpenCV primitives called through the real cv2 and the glue -- layer layout, per-level quotas,
KeyPointsFilter::retainBest (libstdc++ std::nth_element + std::partition), HarrisResponses, ICAngles -- written out.
Checked here against cv2.ORB_create(...).detect / .compute end to end (tests/golden/make_golden_bird.py), which pins
the element ORDER too.  Test infrastructure only."""
import math

import cv2
import numpy as np

F32 = np.float32
HARRIS_K = F32(0.04)


def cv_round(v):
    return int(np.rint(v))


# ---- libstdc++ algorithms on a python list, comp(a, b) = a.response > b.response -------------------------------
def _move_median_to_first(v, result, a, b, c, comp):
    if comp(v[a], v[b]):
        if comp(v[b], v[c]):
            v[result], v[b] = v[b], v[result]
        elif comp(v[a], v[c]):
            v[result], v[c] = v[c], v[result]
        else:
            v[result], v[a] = v[a], v[result]
    elif comp(v[a], v[c]):
        v[result], v[a] = v[a], v[result]
    elif comp(v[b], v[c]):
        v[result], v[c] = v[c], v[result]
    else:
        v[result], v[b] = v[b], v[result]


def _unguarded_partition(v, first, last, pivot, comp):
    while True:
        while comp(v[first], v[pivot]):
            first += 1
        last -= 1
        while comp(v[pivot], v[last]):
            last -= 1
        if not first < last:
            return first
        v[first], v[last] = v[last], v[first]
        first += 1


def _insertion_sort(v, first, last, comp):
    if first == last:
        return
    for i in range(first + 1, last):
        val = v[i]
        if comp(val, v[first]):
            v[first + 1:i + 1] = v[first:i]
            v[first] = val
        else:
            j = i
            while comp(val, v[j - 1]):
                v[j] = v[j - 1]
                j -= 1
            v[j] = val


def nth_element(v, nth, comp):
    first, last = 0, len(v)
    if first == last or nth == last:
        return
    depth = 2 * (len(v).bit_length() - 1)
    while last - first > 3:
        if depth == 0:
            raise NotImplementedError("heap_select fallback of std::nth_element")
        depth -= 1
        mid = first + (last - first) // 2
        _move_median_to_first(v, first, first + 1, mid, last - 1, comp)
        cut = _unguarded_partition(v, first + 1, last, first, comp)
        if cut <= nth:
            first = cut
        else:
            last = cut
    _insertion_sort(v, first, last, comp)


def partition(v, first, last, pred):
    while True:
        while True:
            if first == last:
                return first
            if pred(v[first]):
                first += 1
            else:
                break
        last -= 1
        while True:
            if first == last:
                return first
            if not pred(v[last]):
                last -= 1
            else:
                break
        v[first], v[last] = v[last], v[first]
        first += 1


def retain_best(kps, n_points):
    """KeyPointsFilter::retainBest; kps = list of [x, y, response, ...] (response at index 2)"""
    if n_points >= 0 and len(kps) > n_points:
        if n_points == 0:
            del kps[:]
            return
        nth_element(kps, n_points - 1, lambda a, b: a[2] > b[2])
        amb = kps[n_points - 1][2]
        end = partition(kps, n_points, len(kps), lambda a: a[2] >= amb)
        del kps[end:]


# ---- ORB ----------------------------------------------------------------------------------------------------------
def get_scale(level, first_level, scale_factor):
    return F32(math.pow(float(F32(scale_factor)), float(level - first_level)))


def layout(w, h, nlevels=8, scale_factor=1.2, edge_threshold=31, patch_size=31):
    half = patch_size // 2
    desc_patch = int(math.ceil(half * math.sqrt(2.0)))
    border = max(edge_threshold, desc_patch, 9 // 2) + 1
    inv0 = F32(1.0) / get_scale(0, 0, scale_factor)
    l0w, l0h = cv_round(F32(w) * inv0), cv_round(F32(h) * inv0)
    buf_w = (l0w + border * 2 + 15) // 16 * 16
    level_dy = l0h + border * 2
    ox, oy = 0, 0
    layers, scales = [], []
    for level in range(nlevels):
        scale = get_scale(level, 0, scale_factor)
        inv = F32(1.0) / scale
        sw, sh = cv_round(F32(w) * inv), cv_round(F32(h) * inv)
        ww, wh = sw + border * 2, sh + border * 2
        if ox + ww > buf_w:
            ox, oy = 0, oy + level_dy
            level_dy = wh
        layers.append((ox + border, oy + border, sw, sh))
        scales.append(scale)
        ox += ww
    return border, buf_w, oy + level_dy, layers, scales


def build_pyramid(image, mask, nlevels=8, scale_factor=1.2):
    h, w = image.shape
    border, bw, bh, layers, scales = layout(w, h, nlevels, scale_factor)
    pyr = np.zeros((bh, bw), np.uint8)
    mpyr = np.zeros((bh, bw), np.uint8) if mask is not None else None
    prev, prev_m = image, mask
    for level, (x, y, sw, sh) in enumerate(layers):
        if level != 0:
            cur = cv2.resize(prev, (sw, sh), interpolation=cv2.INTER_LINEAR_EXACT)
            if mask is not None:
                cur_m = cv2.resize(prev_m, (sw, sh), interpolation=cv2.INTER_LINEAR_EXACT)
                _, cur_m = cv2.threshold(cur_m, 254, 0, cv2.THRESH_TOZERO)
        else:
            cur, cur_m = image, mask
        pyr[y - border:y + sh + border, x - border:x + sw + border] = cv2.copyMakeBorder(cur, border, border, border, border,
                                                                                         cv2.BORDER_REFLECT_101)
        if mask is not None:
            mpyr[y - border:y + sh + border, x - border:x + sw + border] = cv2.copyMakeBorder(cur_m, border, border, border, border,
                                                                                              cv2.BORDER_CONSTANT, value=0)
        if level > 0:
            prev, prev_m = cur, (cur_m if mask is not None else None)
    return pyr, mpyr, layers, scales, border


def features_per_level(nfeatures, nlevels=8, scale_factor=1.2):
    factor = F32(1.0 / scale_factor)
    nd = F32(nfeatures) * (F32(1) - factor) / (F32(1) - F32(math.pow(float(factor), float(nlevels))))
    out, s = [], 0
    for _ in range(nlevels - 1):
        out.append(cv_round(nd))
        s += out[-1]
        nd = F32(nd * factor)
    out.append(max(nfeatures - s, 0))
    return out


def umax_table(half=15):
    umax = [0] * (half + 2)
    vmax = int(math.floor(float(F32(half) * F32(math.sqrt(F32(2.0))) / F32(2) + F32(1))))
    vmin = int(math.ceil(float(F32(half) * F32(math.sqrt(F32(2.0))) / F32(2))))
    for v in range(vmax + 1):
        umax[v] = cv_round(math.sqrt(float(half * half - v * v)))
    v0 = 0
    for v in range(half, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax


def harris(pyr, layers, kps, block=7):
    r = block // 2
    scale = F32(1.0) / (F32(4 * block) * F32(255.0))
    ssq = scale * scale * scale * scale
    P = pyr.astype(np.int32)
    for kp in kps:
        x0, y0, z = cv_round(kp[0]), cv_round(kp[1]), kp[4]
        lx, ly = layers[z][0], layers[z][1]
        ys, xs = y0 - r + ly, x0 - r + lx
        W = P[ys - 1:ys + block + 1, xs - 1:xs + block + 1]
        Ix = (W[1:-1, 2:] - W[1:-1, :-2]) * 2 + (W[:-2, 2:] - W[:-2, :-2]) + (W[2:, 2:] - W[2:, :-2])
        Iy = (W[2:, 1:-1] - W[:-2, 1:-1]) * 2 + (W[2:, :-2] - W[:-2, :-2]) + (W[2:, 2:] - W[:-2, 2:])
        a, b, c = int((Ix * Ix).sum()), int((Iy * Iy).sum()), int((Ix * Iy).sum())
        fa, fb, fc = F32(a), F32(b), F32(c)
        kp[2] = F32((F32(fa * fb) - F32(fc * fc) - F32(F32(HARRIS_K * F32(fa + fb)) * F32(fa + fb))) * ssq)


def ic_angles(pyr, layers, kps, umax, half=15):
    for kp in kps:
        lx, ly = layers[kp[4]][0], layers[kp[4]][1]
        cy, cx = cv_round(kp[1]) + ly, cv_round(kp[0]) + lx
        m01 = m10 = 0
        row = pyr[cy].astype(np.int32)
        for u in range(-half, half + 1):
            m10 += u * int(row[cx + u])
        for v in range(1, half + 1):
            d = umax[v]
            p = pyr[cy + v, cx - d:cx + d + 1].astype(np.int32)
            m = pyr[cy - v, cx - d:cx + d + 1].astype(np.int32)
            us = np.arange(-d, d + 1)
            m01 += v * int((p - m).sum())
            m10 += int((us * (p + m)).sum())
        kp[3] = F32(cv2.fastAtan2(float(m01), float(m10)))


def detect(image, mask, nfeatures=2000, nlevels=8, scale_factor=1.2, edge_threshold=31, patch_size=31, fast_threshold=20):
    """-> list of [x, y, response, angle, octave, size] in cv::ORB's output order"""
    pyr, mpyr, layers, scales, border = build_pyramid(image, mask, nlevels, scale_factor)
    quota = features_per_level(nfeatures, nlevels, scale_factor)
    umax = umax_table(patch_size // 2)
    fd = cv2.FastFeatureDetector_create(fast_threshold, True)
    allk, counters = [], []
    for level, (x, y, sw, sh) in enumerate(layers):
        img = np.ascontiguousarray(pyr[y:y + sh, x:x + sw])
        msk = None if mpyr is None else np.ascontiguousarray(mpyr[y:y + sh, x:x + sw])
        kps = [[kp.pt[0], kp.pt[1], F32(kp.response), F32(-1), level, 0.0] for kp in fd.detect(img, None)]
        if msk is not None:                                    # KeyPointsFilter::runByPixelsMask
            kps = [k for k in kps if msk[int(F32(k[1]) + F32(0.5)), int(F32(k[0]) + F32(0.5))] != 0]
        if sh <= edge_threshold * 2 or sw <= edge_threshold * 2:   # runByImageBorder
            kps = []
        else:
            kps = [k for k in kps if edge_threshold <= cv_round(k[0]) < sw - edge_threshold and edge_threshold <= cv_round(k[1]) < sh - edge_threshold]
        retain_best(kps, 2 * quota[level])
        for k in kps:
            k[5] = F32(F32(patch_size) * scales[level])
        counters.append(len(kps))
        allk += kps
    harris(pyr, layers, allk)
    out, off = [], 0
    for level in range(nlevels):
        kps = allk[off:off + counters[level]]
        off += counters[level]
        retain_best(kps, quota[level])
        out += kps
    ic_angles(pyr, layers, out, umax, patch_size // 2)
    for k in out:
        s = scales[k[4]]
        k[0], k[1] = F32(F32(k[0]) * s), F32(F32(k[1]) * s)
    return out, (pyr, layers, scales, border)
