"""Literal Python transcription of Frame::ComputeStereoMatches (reference src/Frame.cc:662-836) on plain arrays:
pins the oracle's C++ restatement (small cases only)."""
import math

import numpy as np

F32 = np.float32
TH_HIGH, TH_LOW = 100, 50


def c_round(v):
    v = float(v)
    return math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)


def dist256(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def compute_stereo_matches(pyr_l, pyr_r, kps_l, desc_l, kps_r, desc_r, scale, inv_scale, mb, mbf):
    N, Nr = len(kps_l), len(kps_r)
    u_right = np.full(N, -1.0, F32)
    depth = np.full(N, -1.0, F32)
    th_orb = (TH_HIGH + TH_LOW) // 2
    n_rows = pyr_l[0].shape[0]
    rows = [[] for _ in range(n_rows)]
    for iR in range(Nr):
        kp_y = F32(kps_r["y"][iR])
        r = F32(F32(2.0) * F32(scale[int(kps_r["octave"][iR])]))
        maxr = int(math.ceil(float(F32(kp_y + r))))
        minr = int(math.floor(float(F32(kp_y - r))))
        for yi in range(minr, maxr + 1):
            rows[yi].append(iR)
    min_z = F32(mb)
    min_d = F32(0)
    max_d = F32(F32(mbf) / min_z)
    dist_idx = []
    for iL in range(N):
        level_l = int(kps_l["octave"][iL])
        v_l, u_l = F32(kps_l["y"][iL]), F32(kps_l["x"][iL])
        cands = rows[int(v_l)]
        if not cands:
            continue
        min_u = F32(u_l - max_d)
        max_u = F32(u_l - min_d)
        if max_u < 0:
            continue
        best, best_r = TH_HIGH, 0
        for iR in cands:
            o = int(kps_r["octave"][iR])
            if o < level_l - 1 or o > level_l + 1:
                continue
            u_r = F32(kps_r["x"][iR])
            if min_u <= u_r <= max_u:
                d = dist256(desc_l[iL], desc_r[iR])
                if d < best:
                    best, best_r = d, iR
        if best < th_orb:
            u_r0 = F32(kps_r["x"][best_r])
            sfac = F32(inv_scale[level_l])
            su_l = F32(c_round(F32(u_l * sfac)))
            sv_l = F32(c_round(F32(v_l * sfac)))
            su_r0 = F32(c_round(F32(u_r0 * sfac)))
            w, L = 5, 5
            PL, PR = pyr_l[level_l], pyr_r[level_l]
            y0, x0 = int(sv_l - w), int(su_l - w)
            IL = PL[y0:y0 + 2 * w + 1, x0:x0 + 2 * w + 1].astype(F32)
            IL = IL - IL[w, w]
            best_s, best_inc = 2 ** 31 - 1, 0
            dists = [F32(0)] * (2 * L + 1)
            iniu = F32(su_r0 + L - w)
            endu = F32(su_r0 + L + w + 1)
            if iniu < 0 or endu >= PR.shape[1]:
                continue
            for inc in range(-L, L + 1):
                xr = int(su_r0 + inc - w)
                IR = PR[y0:y0 + 2 * w + 1, xr:xr + 2 * w + 1].astype(F32)
                IR = IR - IR[w, w]
                d = F32(np.abs(IL - IR).sum(dtype=np.float64))
                if d < best_s:
                    best_s, best_inc = int(d), inc
                dists[L + inc] = d
            if best_inc == -L or best_inc == L:
                continue
            d1, d2, d3 = dists[L + best_inc - 1], dists[L + best_inc], dists[L + best_inc + 1]
            den = F32(F32(2.0) * F32(F32(d1 + d3) - F32(F32(2.0) * d2)))
            with np.errstate(divide="ignore", invalid="ignore"):
                delta = F32(F32(d1 - d3) / den)
            if delta < -1 or delta > 1:
                continue
            best_u = F32(F32(scale[level_l]) * F32(F32(su_r0 + F32(best_inc)) + delta))
            disp = F32(u_l - best_u)
            if disp >= min_d and disp < max_d:
                if disp <= 0:
                    disp = F32(0.01)
                    best_u = F32(float(u_l) - 0.01)
                depth[iL] = F32(F32(mbf) / disp)
                u_right[iL] = best_u
                dist_idx.append((best_s, iL))
    if not dist_idx:
        return 0, u_right, depth
    dist_idx.sort()
    median = F32(dist_idx[len(dist_idx) // 2][0])
    th = F32(F32(F32(1.5) * F32(1.4)) * median)
    kept = len(dist_idx)
    for d, i in reversed(dist_idx):
        if F32(d) < th:
            break
        u_right[i] = -1
        depth[i] = -1
        kept -= 1
    return kept, u_right, depth
