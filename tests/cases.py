"""Seeded synthetic matcher cases shared by the oracle tests (CPU) and the GPU parity tests.

Everything is generated from numpy PCG64 streams; no reference files are read at run time.
"""
import numpy as np

from helpers import KP_DTYPE, synth

QUOTA_FRAC = np.array([217, 181, 151, 126, 105, 87, 73, 60], np.float64) / 1000.0
SCALE_FACTORS = np.array([np.float32(1.2) ** 0] * 8, np.float32)
_s = np.float32(1.0)
for _i in range(8):
    SCALE_FACTORS[_i] = _s
    _s = np.float32(np.float64(_s) * np.float64(np.float32(1.2)))


def random_keypoints(n, w, h, seed, integer_level_coords=True):
    """Keypoints shaped like extractor output: octave histogram from the per-level quota, level-integer
    coordinates scaled by the level's scale factor, random angle/response."""
    rng = np.random.default_rng(seed)
    kps = np.zeros(n, KP_DTYPE)
    octv = rng.choice(8, size=n, p=QUOTA_FRAC / QUOTA_FRAC.sum()).astype(np.int32)
    octv.sort()
    sf = SCALE_FACTORS[octv]
    lx = rng.integers(19, np.maximum(20, (w / sf).astype(np.int64) - 19))
    ly = rng.integers(19, np.maximum(20, (h / sf).astype(np.int64) - 19))
    kps["x"] = (lx.astype(np.float32) * sf).astype(np.float32)
    kps["y"] = (ly.astype(np.float32) * sf).astype(np.float32)
    kps["size"] = (31 * sf).astype(np.int32).astype(np.float32)
    kps["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    kps["response"] = rng.integers(7, 120, n).astype(np.float32)
    kps["octave"] = octv
    kps["class_id"] = -1
    return kps


def frame_case(n, w, h, seed, stereo_frac=0.0):
    """(kps, desc, u_right, grid params) of a synthetic frame"""
    rng = np.random.default_rng(seed + 7)
    kps = random_keypoints(n, w, h, seed)
    desc = synth.synth_descriptors(n, seed + 1)
    u_right = np.full(n, -1, np.float32)
    if stereo_frac > 0:
        m = rng.random(n) < stereo_frac
        u_right[m] = kps["x"][m] - rng.uniform(1, 40, int(m.sum())).astype(np.float32)
    grid = dict(min_x=np.float32(0), min_y=np.float32(0),
                inv_w=np.float32(64) / np.float32(w), inv_h=np.float32(48) / np.float32(h))
    return kps, desc, u_right, grid


def projection_queries(kps, desc, u_right, w, h, nq, seed, on_kp_frac=0.7, max_flips=40, jitter=3.0):
    """Map-point style queries: a fraction sit on (jittered) frame keypoints with a perturbed copy of that
    keypoint's descriptor (so windows are non-empty and ratio tests bite), the rest are uniform random."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    src = rng.integers(0, n, nq)
    on = rng.random(nq) < on_kp_frac
    u = np.where(on, kps["x"][src] + rng.uniform(-jitter, jitter, nq), rng.uniform(0, w, nq)).astype(np.float32)
    v = np.where(on, kps["y"][src] + rng.uniform(-jitter, jitter, nq), rng.uniform(0, h, nq)).astype(np.float32)
    lvl = np.where(on, np.clip(kps["octave"][src] + rng.integers(0, 2, nq), 0, 7), rng.integers(0, 8, nq)).astype(np.int32)
    qd = synth.perturb_descriptors(desc[src], max_flips, seed + 1)
    rnd = synth.synth_descriptors(nq, seed + 2)
    qd = np.where(on[:, None], qd, rnd)
    viewcos = np.where(rng.random(nq) < 0.5, 0.999, 0.9).astype(np.float32)
    uR = np.where(u_right[src] > 0, u_right[src] + rng.uniform(-2, 2, nq), u - 10).astype(np.float32)
    valid = (rng.random(nq) < 0.93).astype(np.uint8)
    obs_pos = (rng.random(nq) < 0.9).astype(np.uint8)
    angle = np.where(on, kps["angle"][src] + rng.normal(0, 4, nq), rng.uniform(0, 360, nq)) % 360
    return dict(valid=valid, u=u, v=v, uR=uR, level=lvl, viewcos=viewcos, desc=qd, obs_pos=obs_pos,
                angle=angle.astype(np.float32), invz=rng.uniform(0.02, 0.5, nq).astype(np.float32), src=src)


def bird_pair(n, size, seed, shift=(3, -2), max_flips=25):
    """Two consecutive birdview frames: frame 2 = frame 1 shifted by <=5 px with perturbed descriptors,
    some points dropped/added."""
    rng = np.random.default_rng(seed)
    k1, d1, _, grid = frame_case(n, size, size, seed)
    keep = rng.random(n) < 0.85
    k2 = k1[keep].copy()
    k2["x"] = k2["x"] + np.float32(shift[0]) + rng.integers(-1, 2, len(k2)).astype(np.float32)
    k2["y"] = k2["y"] + np.float32(shift[1]) + rng.integers(-1, 2, len(k2)).astype(np.float32)
    k2["angle"] = (k2["angle"] + rng.normal(0, 3, len(k2)).astype(np.float32)) % np.float32(360)
    d2 = synth.perturb_descriptors(d1[keep], max_flips, seed + 3)
    extra = int(0.15 * n)
    ke, de, _, _ = frame_case(extra, size, size, seed + 11)
    k2 = np.concatenate([k2, ke])
    d2 = np.concatenate([d2, de])
    perm = rng.permutation(len(k2))
    return (k1, d1), (k2[perm].copy(), d2[perm].copy()), grid


def triangulation_case(n1, n2, w, h, seed, n_nodes=60):
    rng = np.random.default_rng(seed)
    k1, d1, uR1, _ = frame_case(n1, w, h, seed, stereo_frac=0.3)
    src = rng.integers(0, n1, n2)
    k2 = k1[src].copy()
    k2["x"] += rng.uniform(-25, 25, n2).astype(np.float32)
    k2["y"] += rng.uniform(-1.5, 1.5, n2).astype(np.float32)
    k2["angle"] = (k2["angle"] + rng.normal(0, 5, n2).astype(np.float32)) % np.float32(360)
    d2 = synth.perturb_descriptors(d1[src], 30, seed + 5)
    uR2 = np.where(rng.random(n2) < 0.3, k2["x"] - 5, -1).astype(np.float32)
    node1 = rng.integers(0, n_nodes, n1)
    node2 = np.where(rng.random(n2) < 0.8, node1[src], rng.integers(0, n_nodes, n2))
    # drop some node ids from each side so that the lower_bound skips are exercised
    node1 = np.where(node1 % 7 == 3, node1 + n_nodes, node1)
    node2 = np.where(node2 % 11 == 5, node2 + 2 * n_nodes, node2)

    def csr(nodes):
        ids = np.unique(nodes)
        ptr = [0]
        idx = []
        for nid in ids:
            members = np.nonzero(nodes == nid)[0]
            idx.extend(members.tolist())
            ptr.append(len(idx))
        return ids.astype(np.int32), np.array(ptr, np.int32), np.array(idx, np.int32)

    has1 = (rng.random(n1) < 0.3).astype(np.uint8)
    has2 = (rng.random(n2) < 0.3).astype(np.uint8)
    # pure horizontal translation: F12 = [t]_x with t=(1,0,0): epipolar lines are image rows
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)
    sigma2 = (SCALE_FACTORS * SCALE_FACTORS).astype(np.float32)
    return dict(k1=k1, d1=d1, uR1=uR1, has1=has1, k2=k2, d2=d2, uR2=uR2, has2=has2, fv1=csr(node1), fv2=csr(node2),
                F12=F12, ex=np.float32(w * 0.5), ey=np.float32(h * 0.5), sf2=SCALE_FACTORS, sigma2=sigma2)


def csr_from_nodes(nodes):
    """node id per keypoint -> (ascending ids, ptr, idx) = DBoW2::FeatureVector as CSR"""
    ids = np.unique(nodes)
    ptr, idx = [0], []
    for nid in ids:
        idx.extend(np.nonzero(nodes == nid)[0].tolist())
        ptr.append(len(idx))
    return ids.astype(np.int32), np.array(ptr, np.int32), np.array(idx, np.int32)


def csr_to_dict(fv):
    return {int(n): fv[2][fv[1][i]:fv[1][i + 1]].tolist() for i, n in enumerate(fv[0])}


def bow_case(n1, n2, w, h, seed, n_nodes=80, max_flips=25):
    """Two keyframes sharing vocabulary nodes: KF2 keypoints are perturbed copies of KF1 keypoints placed
    (mostly) in the same node, plus distractors; some nodes exist on one side only."""
    rng = np.random.default_rng(seed)
    k1, d1, _, grid = frame_case(n1, w, h, seed)
    src = rng.integers(0, n1, n2)
    k2 = k1[src].copy()
    k2["angle"] = (k2["angle"] + rng.normal(0, 6, n2).astype(np.float32)) % np.float32(360)
    d2 = synth.perturb_descriptors(d1[src], max_flips, seed + 3)
    rnd = rng.random(n2) < 0.25
    d2[rnd] = synth.synth_descriptors(int(rnd.sum()), seed + 4)
    node1 = rng.integers(0, n_nodes, n1)
    node2 = np.where(rng.random(n2) < 0.85, node1[src], rng.integers(0, n_nodes, n2))
    node1 = np.where(node1 % 9 == 4, node1 + n_nodes, node1)
    node2 = np.where(node2 % 13 == 6, node2 + 2 * n_nodes, node2)
    valid1 = (rng.random(n1) < 0.7).astype(np.uint8)
    valid2 = (rng.random(n2) < 0.8).astype(np.uint8)
    return dict(k1=k1, d1=d1, k2=k2, d2=d2, fv1=csr_from_nodes(node1), fv2=csr_from_nodes(node2), valid1=valid1, valid2=valid2, grid=grid)


def best_window_queries(kps, desc, u_right, w, h, nq, seed, th):
    """Projected map points with a predicted level (SearchByProjection(Frame,KF,set) / (KF,Scw) / Fuse style)"""
    q = projection_queries(kps, desc, u_right, w, h, nq, seed, jitter=2.0)
    pred = q["level"]
    r = (np.float32(th) * SCALE_FACTORS[pred]).astype(np.float32)
    ur = (q["u"] - np.float32(12.0) * q["invz"]).astype(np.float32)
    return dict(valid=q["valid"], u=q["u"], v=q["v"], pred=pred, r=r, ur=ur, angle=q["angle"], desc=q["desc"], obs_pos=q["obs_pos"])


def synthetic_vocabulary(k=10, L=3, seed=0, stop_frac=0.03):
    """A DBoW2-style vocabulary tree (k children, L levels) with random node descriptors; children are small
    perturbations of their parent so that the descent is meaningful.  ORBvoc itself (k=10, L=6) is not part of the
    reference checkout (.MISSING_LARGE_BLOBS); the algorithm does not depend on the tree's contents."""
    rng = np.random.default_rng(seed)
    desc = [synth.synth_descriptors(1, seed)[0]]
    children = [[]]
    level_nodes = [0]
    for lvl in range(L):
        nxt = []
        for p in level_nodes:
            for _ in range(k):
                nid = len(desc)
                flips = 64 >> lvl
                d = synth.perturb_descriptors(desc[p][None, :], flips, int(rng.integers(1 << 30)))[0]
                desc.append(d)
                children.append([])
                children[p].append(nid)
                nxt.append(nid)
        level_nodes = nxt
    n = len(desc)
    word_id = np.full(n, -1, np.int32)
    weight = np.zeros(n, np.float64)
    for w, nid in enumerate(level_nodes):
        word_id[nid] = w
        weight[nid] = 0.0 if rng.random() < stop_frac else float(rng.uniform(0.5, 8.0))    # idf; 0 = stopped word
    ptr = np.zeros(n + 1, np.int32)
    idx = []
    for i, c in enumerate(children):
        idx.extend(c)
        ptr[i + 1] = len(idx)
    return dict(child_ptr=ptr, child_idx=np.array(idx, np.int32), node_desc=np.stack(desc), word_id=word_id, weight=weight, L=L)


def local_map_case(kps, desc, w, h, n, seed, on_kp_frac=0.6, max_flips=40):
    """A local map + camera pose for Frame::isInFrustum / SearchLocalPoints: a fraction of the map points are
    back-projections of frame keypoints at random depths (their descriptors perturbed copies), the rest are random
    world points, some behind the camera, outside the image, out of their scale-invariance range or seen from a
    bad angle.  Returns (pose dict, map dict)."""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    fx, fy, cx, cy, bf = f32(718.856), f32(718.856), f32(w / 2 - 13.3), f32(h / 2 - 2.8), f32(386.1448)
    ax, ay, az = rng.normal(0, 0.05, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    R = (Rz @ Ry @ Rx).astype(f32)
    t = rng.normal(0, 2.0, 3).astype(f32)
    Ow = (-(R.T.astype(f32) @ t)).astype(f32)
    nk = len(kps)
    src = rng.integers(0, nk, n)
    on = rng.random(n) < on_kp_frac
    z = rng.uniform(2.0, 60.0, n)
    uu = np.where(on, kps["x"][src] + rng.uniform(-2, 2, n), rng.uniform(-0.2 * w, 1.2 * w, n))
    vv = np.where(on, kps["y"][src] + rng.uniform(-2, 2, n), rng.uniform(-0.2 * h, 1.2 * h, n))
    z = np.where(~on & (rng.random(n) < 0.15), -z, z)                       # some behind the camera
    Xc = np.stack([(uu - cx) * z / fx, (vv - cy) * z / fy, z], 1)
    P = ((Xc - t[None, :].astype(np.float64)) @ R.astype(np.float64)).astype(f32)          # R^T (Xc - t)
    PO = P.astype(np.float64) - Ow[None, :]
    d = np.linalg.norm(PO, axis=1)
    nrm = PO / np.maximum(d, 1e-9)[:, None]
    tilt = rng.normal(0, 0.5, (n, 3)) * (rng.random(n) < 0.3)[:, None] * 3.0
    nrm = nrm + tilt
    nrm = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(f32)
    lvl = np.where(on, kps["octave"][src], rng.integers(0, 8, n))
    maxd = (d * np.power(1.2, lvl) * rng.uniform(0.9, 1.1, n)).astype(f32)
    far = rng.random(n) < 0.1
    maxd = np.where(far, maxd * rng.choice([0.2, 6.0], n), maxd).astype(f32)    # out of range either way
    mind = (maxd / f32(1.2 ** 7)).astype(f32)
    qd = np.where(on[:, None], synth.perturb_descriptors(desc[src], max_flips, seed + 1), synth.synth_descriptors(n, seed + 2))
    pose = dict(Rcw=R.reshape(9), tcw=t, Ow=Ow, fx=float(fx), fy=float(fy), cx=float(cx), cy=float(cy), mbf=float(bf),
                min_x=0.0, max_x=float(w), min_y=0.0, max_y=float(h), log_scale_factor=float(np.log(f32(1.2), dtype=f32)),
                n_levels=8)
    mp = dict(pos=P, normal=nrm, max_distance=maxd, min_distance=mind, desc=np.ascontiguousarray(qd, np.uint8),
              candidate=(rng.random(n) < 0.95).astype(np.uint8), obs_pos=(rng.random(n) < 0.9).astype(np.uint8))
    return pose, mp


def birdview_case(size, seed, vehicle=(80, 140)):
    """Birdview image + mask as src/Frame.cc:317-327 prepares them: mask 255 with the vehicle footprint (plus the
    15 px boundary) zeroed in the middle."""
    h = w = size if isinstance(size, int) else None
    if h is None:
        w, h = size
    img = synth.synth_frame(h, w, seed)
    mask = np.full((h, w), 255, np.uint8)
    vw, vh = vehicle
    mask[h // 2 - vh // 2:h // 2 + vh // 2 + 1, w // 2 - vw // 2:w // 2 + vw // 2 + 1] = 0
    return img, mask


def distinctive_groups(sizes, seed):
    """Observation descriptors of landmarks for ComputeDistinctiveDescriptors: per landmark a base descriptor and
    noisy copies of it (0-60 flipped bits), a few exact duplicates so that medians tie between rows.
    Returns (desc [total][32], group_ptr [len(sizes)+1])."""
    rng = np.random.default_rng(seed)
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    desc = np.zeros((int(ptr[-1]), 32), np.uint8)
    for g, n in enumerate(sizes):
        if n == 0:
            continue
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        bits = np.unpackbits(np.tile(base, (n, 1)), axis=1)
        for i in range(n):
            flips = rng.choice(256, size=int(rng.integers(0, 61)), replace=False)
            bits[i, flips] ^= 1
        rows = np.packbits(bits, axis=1)
        if n >= 4:
            rows[n - 1] = rows[0]          # duplicates: equal rows of the distance matrix -> first index must win
            rows[n // 2] = rows[1]
        desc[ptr[g]:ptr[g + 1]] = rows
    return desc, ptr
