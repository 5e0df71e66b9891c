"""Literal Python transcription of the ORBmatcher Hamming scans and the Frame lookup grid
(reference src/ORBmatcher.cc, src/Frame.cc:378-412,494-559,879-944) on flattened arrays.

Second, independent restatement used to pin oracle/match_oracle.cpp (pure-Python loops: small cases only).
Float expressions use numpy float32 scalars so that they round like the reference's `float` code.
"""
import math

import numpy as np

F32 = np.float32
TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30
GRID_COLS, GRID_ROWS = 64, 48
INT_MAX = 2 ** 31 - 1


def c_round(v):
    """C round(): half away from zero"""
    v = float(v)
    return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))


def descriptor_distance(a, b):
    pa = np.frombuffer(np.ascontiguousarray(a).tobytes(), dtype="<u4")
    pb = np.frombuffer(np.ascontiguousarray(b).tobytes(), dtype="<u4")
    dist = 0
    for i in range(8):
        v = int(pa[i]) ^ int(pb[i])
        v = v - ((v >> 1) & 0x55555555)
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333)
        dist += ((((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) & 0xFFFFFFFF) >> 24
    return dist


class PyFrame:
    def __init__(self, kps, desc, min_x, min_y, inv_w, inv_h, u_right=None):
        self.kps, self.desc = kps, desc
        self.n = len(kps)
        self.min_x, self.min_y, self.inv_w, self.inv_h = F32(min_x), F32(min_y), F32(inv_w), F32(inv_h)
        self.u_right = np.full(self.n, -1, F32) if u_right is None else np.asarray(u_right, F32)
        self.grid = [[[] for _ in range(GRID_ROWS)] for _ in range(GRID_COLS)]
        for i in range(self.n):
            px = c_round((F32(kps["x"][i]) - self.min_x) * self.inv_w)
            py = c_round((F32(kps["y"][i]) - self.min_y) * self.inv_h)
            if px < 0 or px >= GRID_COLS or py < 0 or py >= GRID_ROWS:
                continue
            self.grid[px][py].append(i)

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        x, y, r = F32(x), F32(y), F32(r)
        out = []
        x0 = max(0, int(math.floor(float((x - self.min_x - r) * self.inv_w))))
        if x0 >= GRID_COLS:
            return out
        x1 = min(GRID_COLS - 1, int(math.ceil(float((x - self.min_x + r) * self.inv_w))))
        if x1 < 0:
            return out
        y0 = max(0, int(math.floor(float((y - self.min_y - r) * self.inv_h))))
        if y0 >= GRID_ROWS:
            return out
        y1 = min(GRID_ROWS - 1, int(math.ceil(float((y - self.min_y + r) * self.inv_h))))
        if y1 < 0:
            return out
        check = (min_level > 0) or (max_level >= 0)
        for ix in range(x0, x1 + 1):
            for iy in range(y0, y1 + 1):
                for j in self.grid[ix][iy]:
                    if check:
                        if self.kps["octave"][j] < min_level:
                            continue
                        if max_level >= 0 and self.kps["octave"][j] > max_level:
                            continue
                    dx = F32(self.kps["x"][j]) - x
                    dy = F32(self.kps["y"][j]) - y
                    if abs(dx) < r and abs(dy) < r:
                        out.append(j)
        return out


def three_maxima(histo):
    max1 = max2 = max3 = 0
    ind1 = ind2 = ind3 = -1
    for i, h in enumerate(histo):
        s = len(h)
        if s > max1:
            max3, max2, max1 = max2, max1, s
            ind3, ind2, ind1 = ind2, ind1, i
        elif s > max2:
            max3, max2 = max2, s
            ind3, ind2 = ind2, i
        elif s > max3:
            max3, ind3 = s, i
    if F32(max2) < F32(0.1) * F32(max1):
        ind2 = ind3 = -1
    elif F32(max3) < F32(0.1) * F32(max1):
        ind3 = -1
    return ind1, ind2, ind3


def rot_bin(a1, a2):
    rot = F32(a1) - F32(a2)
    if rot < 0.0:
        rot = F32(rot + F32(360.0))
    b = c_round(F32(rot * F32(F32(1.0) / F32(HISTO_LENGTH))))
    return 0 if b == HISTO_LENGTH else b


def search_by_projection(F, sf, q_valid, q_u, q_v, q_uR, q_level, q_viewcos, q_desc, q_obs_pos, kp_blocked, th, nnratio):
    blocked = [bool(b) for b in kp_blocked] if kp_blocked is not None else [False] * F.n
    query_of_kp = [-1] * F.n
    nm = 0
    for i in range(len(q_u)):
        if not q_valid[i]:
            continue
        lvl = int(q_level[i])
        r = F32(2.5) if float(q_viewcos[i]) > 0.998 else F32(4.0)
        if th != 1.0:
            r = F32(r * F32(th))
        rad = F32(r * F32(sf[lvl]))
        idxs = F.features_in_area(q_u[i], q_v[i], rad, lvl - 1, lvl)
        if not idxs:
            continue
        bd, bl, bd2, bl2, bi = 256, -1, 256, -1, -1
        for idx in idxs:
            if blocked[idx]:
                continue
            if F.u_right[idx] > 0:
                er = abs(F32(q_uR[i]) - F.u_right[idx])
                if er > rad:
                    continue
            d = descriptor_distance(q_desc[i], F.desc[idx])
            if d < bd:
                bd2, bd, bl2, bl, bi = bd, d, bl, int(F.kps["octave"][idx]), idx
            elif d < bd2:
                bl2, bd2 = int(F.kps["octave"][idx]), d
        if bd <= TH_HIGH:
            if bl == bl2 and F32(bd) > F32(nnratio) * F32(bd2):
                continue
            query_of_kp[bi] = i
            blocked[bi] = bool(q_obs_pos[i]) if q_obs_pos is not None else True
            nm += 1
    return nm, query_of_kp


def search_by_projection_frame(Cur, sf, q_valid, q_u, q_v, q_invz, q_octave, q_angle, q_desc, q_obs_pos, kp_blocked,
                               th, mbf, mode, check_ori):
    blocked = [bool(b) for b in kp_blocked] if kp_blocked is not None else [False] * Cur.n
    query_of_kp = [-1] * Cur.n
    hist = [[] for _ in range(HISTO_LENGTH)]
    nm = 0
    for i in range(len(q_u)):
        if not q_valid[i]:
            continue
        o = int(q_octave[i])
        radius = F32(F32(th) * F32(sf[o]))
        if mode == 1:
            idxs = Cur.features_in_area(q_u[i], q_v[i], radius, o, -1)
        elif mode == 2:
            idxs = Cur.features_in_area(q_u[i], q_v[i], radius, 0, o)
        else:
            idxs = Cur.features_in_area(q_u[i], q_v[i], radius, o - 1, o + 1)
        if not idxs:
            continue
        bd, bi = 256, -1
        for i2 in idxs:
            if blocked[i2]:
                continue
            if Cur.u_right[i2] > 0:
                ur = F32(F32(q_u[i]) - F32(F32(mbf) * F32(q_invz[i])))
                if abs(F32(ur - Cur.u_right[i2])) > radius:
                    continue
            d = descriptor_distance(q_desc[i], Cur.desc[i2])
            if d < bd:
                bd, bi = d, i2
        if bd <= TH_HIGH:
            query_of_kp[bi] = i
            blocked[bi] = bool(q_obs_pos[i]) if q_obs_pos is not None else True
            nm += 1
            if check_ori:
                hist[rot_bin(q_angle[i], Cur.kps["angle"][bi])].append(bi)
    if check_ori:
        keep = three_maxima(hist)
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for j in hist[b]:
                    query_of_kp[j] = -1
                    nm -= 1
    return nm, query_of_kp


def birdview_match(kps1, desc1, F2, prev_xy, window, nnratio, check_ori):
    n1 = len(kps1)
    m12 = [-1] * n1
    m21 = [-1] * F2.n
    md = [INT_MAX] * F2.n
    hist = [[] for _ in range(HISTO_LENGTH)]
    nm = 0
    prev = None if prev_xy is None else np.array(prev_xy, F32).reshape(-1, 2)
    for i1 in range(n1):
        l1 = int(kps1["octave"][i1])
        if prev is not None:
            if l1 > 0:
                continue
            idxs = F2.features_in_area(prev[i1, 0], prev[i1, 1], window, l1, l1)
        else:
            idxs = F2.features_in_area(kps1["x"][i1], kps1["y"][i1], window, l1, l1)
        if not idxs:
            continue
        bd, bd2, bi = INT_MAX, INT_MAX, -1
        for i2 in idxs:
            d = descriptor_distance(desc1[i1], F2.desc[i2])
            if md[i2] <= d:
                continue
            if d < bd:
                bd2, bd, bi = bd, d, i2
            elif d < bd2:
                bd2 = d
        if bd <= TH_LOW and F32(bd) < F32(F32(bd2) * F32(nnratio)):
            if m21[bi] >= 0:
                m12[m21[bi]] = -1
                nm -= 1
            m12[i1] = bi
            m21[bi] = i1
            md[bi] = bd
            nm += 1
            if check_ori:
                hist[rot_bin(kps1["angle"][i1], F2.kps["angle"][bi])].append(i1)
    if check_ori:
        keep = three_maxima(hist)
        for b in range(HISTO_LENGTH):
            if b in keep:
                continue
            for i1 in hist[b]:
                if m12[i1] >= 0:
                    m21[m12[i1]] = -1
                    m12[i1] = -1
                    nm -= 1
    if prev is not None:
        for i1 in range(n1):
            if m12[i1] >= 0:
                prev[i1, 0] = F2.kps["x"][m12[i1]]
                prev[i1, 1] = F2.kps["y"][m12[i1]]
    return nm, m12, prev


def search_by_match_bird_kf(kf_kps, has_mp, mp_desc, F, r, nnratio, check_ori):
    out = [-1] * F.n
    md = [INT_MAX] * F.n
    hist = [[] for _ in range(HISTO_LENGTH)]
    nm = 0
    for k in range(len(kf_kps)):
        if not has_mp[k]:
            continue
        idxs = F.features_in_area(kf_kps["x"][k], kf_kps["y"][k], r)
        if not idxs:
            continue
        bd, bl, bd2, bl2, bi = INT_MAX, -1, INT_MAX, -1, -1
        for idx in idxs:
            d = descriptor_distance(mp_desc[k], F.desc[idx])
            if md[idx] <= d:
                continue
            if d < bd:
                bd2, bd, bl2, bl, bi = bd, d, bl, int(F.kps["octave"][idx]), idx
            elif d < bd2:
                bl2, bd2 = int(F.kps["octave"][idx]), d
        if bd <= TH_HIGH and (bl != bl2 or F32(bd) < F32(F32(bd2) * F32(nnratio))):
            out[bi] = k
            md[bi] = bd
            if check_ori:
                hist[rot_bin(kf_kps["angle"][k], F.kps["angle"][bi])].append(bi)
            nm += 1
    if check_ori:
        keep = three_maxima(hist)
        for b in range(HISTO_LENGTH):
            if b in keep:
                continue
            for j in hist[b]:
                out[j] = -1
                nm -= 1
    return nm, out


def search_by_projection_bird(F, q_valid, q_x, q_y, q_desc, q_obs_pos, kp_blocked, r, nnratio):
    blocked = [bool(b) for b in kp_blocked] if kp_blocked is not None else [False] * F.n
    out = [-1] * F.n
    nm = 0
    for i in range(len(q_x)):
        if not q_valid[i]:
            continue
        idxs = F.features_in_area(q_x[i], q_y[i], r)
        if not idxs:
            continue
        bd, bl, bd2, bl2, bi = 256, -1, 256, -1, -1
        for idx in idxs:
            if blocked[idx]:
                continue
            d = descriptor_distance(q_desc[i], F.desc[idx])
            if d < bd:
                bd2, bd, bl2, bl, bi = bd, d, bl, int(F.kps["octave"][idx]), idx
            elif d < bd2:
                bl2, bd2 = int(F.kps["octave"][idx]), d
        if bd <= TH_HIGH:
            if bl == bl2 and F32(bd) > F32(nnratio) * F32(bd2):
                continue
            out[bi] = i
            blocked[bi] = bool(q_obs_pos[i]) if q_obs_pos is not None else True
            nm += 1
    return nm, out


def check_epipolar(kp1, kp2, F12, sigma2):
    x1, y1, x2, y2 = F32(kp1["x"]), F32(kp1["y"]), F32(kp2["x"]), F32(kp2["y"])
    F12 = np.asarray(F12, F32).reshape(3, 3)
    a = F32(F32(x1 * F12[0, 0] + y1 * F12[1, 0]) + F12[2, 0])
    b = F32(F32(x1 * F12[0, 1] + y1 * F12[1, 1]) + F12[2, 1])
    c = F32(F32(x1 * F12[0, 2] + y1 * F12[1, 2]) + F12[2, 2])
    num = F32(F32(a * x2 + b * y2) + c)
    den = F32(a * a + b * b)
    if den == 0:
        return False
    dsqr = F32(F32(num * num) / den)
    return float(dsqr) < 3.84 * float(sigma2[int(kp2["octave"])])


def search_for_triangulation(kps1, desc1, uR1, has_mp1, kps2, desc2, uR2, has_mp2, fv1, fv2, F12, ex, ey, sf2, sigma2,
                             only_stereo, check_ori):
    """fv: dict node -> list of keypoint indices (std::map iteration = ascending node id)"""
    m12 = [-1] * len(kps1)
    hist = [[] for _ in range(HISTO_LENGTH)]
    nm = 0
    for node in sorted(set(fv1) & set(fv2)):
        for idx1 in fv1[node]:
            if has_mp1[idx1]:
                continue
            st1 = uR1[idx1] >= 0
            if only_stereo and not st1:
                continue
            bd, bi = TH_LOW, -1
            for idx2 in fv2[node]:
                if has_mp2[idx2]:
                    continue
                st2 = uR2[idx2] >= 0
                if only_stereo and not st2:
                    continue
                d = descriptor_distance(desc1[idx1], desc2[idx2])
                if d > TH_LOW or d > bd:
                    continue
                if not st1 and not st2:
                    dx = F32(F32(ex) - F32(kps2["x"][idx2]))
                    dy = F32(F32(ey) - F32(kps2["y"][idx2]))
                    if F32(dx * dx + dy * dy) < F32(F32(100) * F32(sf2[int(kps2["octave"][idx2])])):
                        continue
                if check_epipolar(kps1[idx1], kps2[idx2], F12, sigma2):
                    bi, bd = idx2, d
            if bi >= 0:
                m12[idx1] = bi
                nm += 1
                if check_ori:
                    hist[rot_bin(kps1["angle"][idx1], kps2["angle"][bi])].append(idx1)
    if check_ori:
        keep = three_maxima(hist)
        for b in range(HISTO_LENGTH):
            if b in keep:
                continue
            for j in hist[b]:
                m12[j] = -1
                nm -= 1
    pairs = [(i, m12[i]) for i in range(len(kps1)) if m12[i] >= 0]
    return nm, pairs


# ---- second batch: literal transcriptions of the remaining ORBmatcher scans (after the host-side geometry) ----
def search_for_initialization(kps1, desc1, F2, prev_xy, window, nnratio, check_ori):
    """src/ORBmatcher.cc:405-520"""
    n1 = len(kps1)
    m12 = [-1] * n1
    m21 = [-1] * F2.n
    md = [INT_MAX] * F2.n
    hist = [[] for _ in range(HISTO_LENGTH)]
    nm = 0
    prev = np.array(prev_xy, F32).reshape(-1, 2)
    for i1 in range(n1):
        l1 = int(kps1["octave"][i1])
        if l1 > 0:
            continue
        idxs = F2.features_in_area(prev[i1, 0], prev[i1, 1], window, l1, l1)
        if not idxs:
            continue
        bd, bd2, bi = INT_MAX, INT_MAX, -1
        for i2 in idxs:
            d = descriptor_distance(desc1[i1], F2.desc[i2])
            if md[i2] <= d:
                continue
            if d < bd:
                bd2, bd, bi = bd, d, i2
            elif d < bd2:
                bd2 = d
        if bd <= TH_LOW and F32(bd) < F32(F32(bd2) * F32(nnratio)):
            if m21[bi] >= 0:
                m12[m21[bi]] = -1
                nm -= 1
            m12[i1] = bi
            m21[bi] = i1
            md[bi] = bd
            nm += 1
            if check_ori:
                hist[rot_bin(kps1["angle"][i1], F2.kps["angle"][bi])].append(i1)
    if check_ori:
        keep = three_maxima(hist)
        for b in range(HISTO_LENGTH):
            if b in keep:
                continue
            for i1 in hist[b]:
                if m12[i1] >= 0:
                    m12[i1] = -1
                    nm -= 1
    for i1 in range(n1):
        if m12[i1] >= 0:
            prev[i1, 0] = F2.kps["x"][m12[i1]]
            prev[i1, 1] = F2.kps["y"][m12[i1]]
    return nm, m12, prev


def search_by_projection_kf(Cur, sf, q_valid, q_u, q_v, q_pred, q_angle, q_desc, kp_has_mp, th, orb_dist, check_ori):
    """SearchByProjection(Frame&, KeyFrame*, set, th, ORBdist), src/ORBmatcher.cc:1472-1599, after projection"""
    has = [bool(b) for b in kp_has_mp]
    out = [-1] * Cur.n
    hist = [[] for _ in range(HISTO_LENGTH)]
    nm = 0
    for i in range(len(q_u)):
        if not q_valid[i]:
            continue
        p = int(q_pred[i])
        radius = F32(F32(th) * F32(sf[p]))
        idxs = Cur.features_in_area(q_u[i], q_v[i], radius, p - 1, p + 1)
        if not idxs:
            continue
        bd, bi = 256, -1
        for i2 in idxs:
            if has[i2]:
                continue
            d = descriptor_distance(q_desc[i], Cur.desc[i2])
            if d < bd:
                bd, bi = d, i2
        if bd <= orb_dist:
            out[bi] = i
            has[bi] = True
            nm += 1
            if check_ori:
                hist[rot_bin(q_angle[i], Cur.kps["angle"][bi])].append(bi)
    if check_ori:
        keep = three_maxima(hist)
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for j in hist[b]:
                    out[j] = -1
                    nm -= 1
    return nm, out


def search_by_projection_scw(KF, sf, q_valid, q_u, q_v, q_pred, q_desc, kp_matched, th):
    """SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th), src/ORBmatcher.cc:290-403, after projection.
    KeyFrame::GetFeaturesInArea has no level argument (src/KeyFrame.cc:586-625); the level test is in the loop."""
    matched = [bool(b) for b in kp_matched]
    out = [-1] * KF.n
    nm = 0
    for i in range(len(q_u)):
        if not q_valid[i]:
            continue
        p = int(q_pred[i])
        radius = F32(F32(th) * F32(sf[p]))
        idxs = KF.features_in_area(q_u[i], q_v[i], radius)
        if not idxs:
            continue
        bd, bi = 256, -1
        for idx in idxs:
            if matched[idx]:
                continue
            lvl = int(KF.kps["octave"][idx])
            if lvl < p - 1 or lvl > p:
                continue
            d = descriptor_distance(q_desc[i], KF.desc[idx])
            if d < bd:
                bd, bi = d, idx
        if bd <= TH_LOW:
            out[bi] = i
            matched[bi] = True
            nm += 1
    return nm, out


def fuse_best(KF, sf, inv_sigma2, q_valid, q_u, q_v, q_ur, q_pred, q_desc, th):
    """matching part of Fuse(KeyFrame*, vpMapPoints, th), src/ORBmatcher.cc:825-975: best keypoint per point"""
    best_idx = [-1] * len(q_u)
    nf = 0
    for i in range(len(q_u)):
        if not q_valid[i]:
            continue
        p = int(q_pred[i])
        u, v, ur = F32(q_u[i]), F32(q_v[i]), F32(q_ur[i])
        radius = F32(F32(th) * F32(sf[p]))
        idxs = KF.features_in_area(u, v, radius)
        if not idxs:
            continue
        bd, bi = 256, -1
        for idx in idxs:
            lvl = int(KF.kps["octave"][idx])
            if lvl < p - 1 or lvl > p:
                continue
            kpx, kpy = F32(KF.kps["x"][idx]), F32(KF.kps["y"][idx])
            if KF.u_right[idx] >= 0:
                ex, ey, er = F32(u - kpx), F32(v - kpy), F32(ur - KF.u_right[idx])
                e2 = F32(F32(F32(ex * ex) + F32(ey * ey)) + F32(er * er))
                if float(F32(e2 * F32(inv_sigma2[lvl]))) > 7.8:
                    continue
            else:
                ex, ey = F32(u - kpx), F32(v - kpy)
                e2 = F32(F32(ex * ex) + F32(ey * ey))
                if float(F32(e2 * F32(inv_sigma2[lvl]))) > 5.99:
                    continue
            d = descriptor_distance(q_desc[i], KF.desc[idx])
            if d < bd:
                bd, bi = d, idx
        if bd <= TH_LOW:
            best_idx[i] = bi
            nf += 1
    return nf, best_idx


def search_by_bow(desc1, angle1, valid1, F2, valid2, fv1, fv2, nnratio, check_ori, kf_kf):
    """SearchByBoW(KeyFrame*, Frame&, ..) src/ORBmatcher.cc:159-288 (kf_kf False) and
    SearchByBoW(KeyFrame*, KeyFrame*, ..) :522-655 (kf_kf True).  fv: dict node -> index list."""
    n1 = len(desc1)
    out_f = [-1] * F2.n
    out_12 = [-1] * n1
    matched2 = [False] * F2.n
    hist = [[] for _ in range(HISTO_LENGTH)]
    nm = 0
    for node in sorted(set(fv1) & set(fv2)):
        for idx1 in fv1[node]:
            if not valid1[idx1]:
                continue
            b1, b2, bi = 256, 256, -1
            for idx2 in fv2[node]:
                if kf_kf:
                    if matched2[idx2] or not valid2[idx2]:
                        continue
                else:
                    if out_f[idx2] >= 0:
                        continue
                d = descriptor_distance(desc1[idx1], F2.desc[idx2])
                if d < b1:
                    b2, b1, bi = b1, d, idx2
                elif d < b2:
                    b2 = d
            ok = (b1 < TH_LOW) if kf_kf else (b1 <= TH_LOW)
            if ok and F32(b1) < F32(F32(nnratio) * F32(b2)):
                if kf_kf:
                    out_12[idx1] = bi
                    matched2[bi] = True
                else:
                    out_f[bi] = idx1
                if check_ori:
                    hist[rot_bin(angle1[idx1], F2.kps["angle"][bi])].append(idx1 if kf_kf else bi)
                nm += 1
    if check_ori:
        keep = three_maxima(hist)
        for b in range(HISTO_LENGTH):
            if b in keep:
                continue
            for j in hist[b]:
                if kf_kf:
                    out_12[j] = -1
                else:
                    out_f[j] = -1
                nm -= 1
    return nm, (out_12 if kf_kf else out_f)


def compute_distinctive_descriptors(vDescriptors):
    """Literal transcription of the selection in MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:272-301;
    identical in MapPointBird.cc:117-146).  vDescriptors: list of 32-byte rows.  Returns (BestIdx, BestMedian) or
    (-1, -1) when the list is empty (the reference returns before choosing)."""
    N = len(vDescriptors)
    if N == 0:
        return -1, -1
    Distances = [[0.0] * N for _ in range(N)]
    for i in range(N):
        Distances[i][i] = 0
        for j in range(i + 1, N):
            distij = int(np.unpackbits(np.bitwise_xor(vDescriptors[i], vDescriptors[j])).sum())
            Distances[i][j] = distij
            Distances[j][i] = distij
    BestMedian = 2 ** 31 - 1
    BestIdx = 0
    for i in range(N):
        vDists = sorted(int(d) for d in Distances[i])
        median = vDists[int(0.5 * (N - 1))]
        if median < BestMedian:
            BestMedian = median
            BestIdx = i
    return BestIdx, BestMedian
