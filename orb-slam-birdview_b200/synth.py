"""Deterministic synthetic inputs for tests and bench.

Frames: uniform noise -> Gaussian blur (sigma 2) -> min-max normalise -> W*H/4000 random rectangles at
50 % alpha (multi-scale corners) -> one low-contrast horizontal band (20 % of the rows, contrast x0.1) so
that the minThFAST fallback and empty FAST cells occur.  numpy only; seed = config_id*1000 + frame_idx.
"""
import numpy as np


def _gauss_sep(a, sigma=2.0, radius=6):
    k = np.exp(-0.5 * (np.arange(-radius, radius + 1) / sigma) ** 2).astype(np.float32)
    k /= k.sum()
    for axis in (0, 1):
        pad = [(0, 0), (0, 0)]
        pad[axis] = (radius, radius)
        p = np.pad(a, pad, mode="reflect")
        out = np.zeros_like(a)
        for i, kv in enumerate(k):
            sl = [slice(None), slice(None)]
            sl[axis] = slice(i, i + a.shape[axis])
            out += kv * p[tuple(sl)]
        a = out
    return a


def synth_frame(h, w, seed, band=True):
    rng = np.random.default_rng(int(seed))
    a = rng.integers(0, 256, size=(h, w), dtype=np.uint8).astype(np.float32)
    a = _gauss_sep(a)
    a = (a - a.min()) * (255.0 / max(float(a.max() - a.min()), 1e-6))
    nrect = max(1, (w * h) // 4000)
    xs = rng.integers(0, w, nrect)
    ys = rng.integers(0, h, nrect)
    ws = rng.integers(8, 61, nrect)
    hs = rng.integers(8, 61, nrect)
    gs = rng.integers(0, 256, nrect)
    for x, y, rw, rh, g in zip(xs, ys, ws, hs, gs):
        a[y:y + rh, x:x + rw] = 0.5 * a[y:y + rh, x:x + rw] + 0.5 * float(g)
    if band:
        bh = max(1, h // 5)
        y0 = int(rng.integers(0, h - bh + 1))
        strip = a[y0:y0 + bh]
        a[y0:y0 + bh] = 110.0 + 0.1 * (strip - 110.0)
    return np.clip(np.rint(a), 0, 255).astype(np.uint8)


def shift_frame(img, dx, dy):
    """Translate by integer (dx, dy) with edge replication (consecutive birdview frames)."""
    h, w = img.shape
    ys = np.clip(np.arange(h) - dy, 0, h - 1)
    xs = np.clip(np.arange(w) - dx, 0, w - 1)
    return np.ascontiguousarray(img[np.ix_(ys, xs)])


def synth_descriptors(n, seed):
    rng = np.random.default_rng(int(seed))
    return rng.integers(0, 256, size=(n, 32), dtype=np.uint8)


def perturb_descriptors(desc, max_flips, seed):
    """Copy of desc with 0..max_flips random bit flips per row."""
    rng = np.random.default_rng(int(seed))
    out = desc.copy()
    bits = np.unpackbits(out, axis=1)
    nflip = rng.integers(0, max_flips + 1, len(out))
    for i, k in enumerate(nflip):
        if k:
            pos = rng.choice(256, size=int(k), replace=False)
            bits[i, pos] ^= 1
    return np.packbits(bits, axis=1)


def projection_queries(kps, desc, w, h, nq, seed, levels=8, on_kp_frac=0.7, max_flips=40, jitter=3.0):
    """Local-map style queries for SearchByProjection: 70 % sit on a
    (jittered) frame keypoint with that keypoint's descriptor under 0..40 random bit flips, 30 % are uniform
    random positions with random descriptors; level from the keypoint octaves; viewCos in {0.999, 0.9}."""
    rng = np.random.default_rng(int(seed))
    n = len(kps)
    src = rng.integers(0, max(n, 1), nq)
    on = (rng.random(nq) < on_kp_frac) & (n > 0)
    kx = kps["x"][src] if n else np.zeros(nq, np.float32)
    ky = kps["y"][src] if n else np.zeros(nq, np.float32)
    ko = kps["octave"][src] if n else np.zeros(nq, np.int32)
    u = np.where(on, kx + rng.uniform(-jitter, jitter, nq), rng.uniform(0, w, nq)).astype(np.float32)
    v = np.where(on, ky + rng.uniform(-jitter, jitter, nq), rng.uniform(0, h, nq)).astype(np.float32)
    lvl = np.where(on, np.clip(ko + rng.integers(0, 2, nq), 0, levels - 1), rng.integers(0, levels, nq)).astype(np.int32)
    base = desc[src] if n else np.zeros((nq, 32), np.uint8)
    # bit flips: XOR with a random mask of up to max_flips set bits (vectorised)
    nflip = rng.integers(0, max_flips + 1, nq)
    bits = (rng.random((nq, 256)).argsort(1) < nflip[:, None]).astype(np.uint8)
    qd = base ^ np.packbits(bits, axis=1)
    rnd = rng.integers(0, 256, size=(nq, 32), dtype=np.uint8)
    qd = np.where(on[:, None], qd, rnd).astype(np.uint8)
    viewcos = np.where(rng.random(nq) < 0.5, 0.999, 0.9).astype(np.float32)
    return dict(valid=np.ones(nq, np.uint8), u=u, v=v, uR=np.full(nq, -1, np.float32), level=lvl, viewcos=viewcos,
                desc=np.ascontiguousarray(qd), obs_pos=np.ones(nq, np.uint8))


# ---------------------------------------------------------------------------------------------------------
# The north-star frame: KITTI-shape stereo pair + birdview image + local map + pose, as a SEQUENCE
# ---------------------------------------------------------------------------------------------------------
KITTI = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, bf=386.1448)       # Examples/Stereo/KITTI00-02.yaml
DISPARITY = 7                                                                     # right view = left shifted by 7 px


def northstar_sequence(n_frames, seed, w=1241, h=376, bird=(400, 400), max_step=5, vehicle=(80, 140)):
    """Images of n_frames consecutive frames.  The front camera translates sideways in front of a fronto-parallel scene
    (depth Z0 = bf / 7 px), so frame i is the window [ox_i, ox_i + w) of one wide canvas and the right image is the left
    one shifted by the 7 px disparity; the birdview camera random-walks (<= max_step px per frame) over a ground canvas.
    Returns dict(imgs [2n][h][w] L/R interleaved, bird_imgs [n][bh][bw], bird_mask [bh][bw], ox [n], poses [n] dicts)."""
    rng = np.random.default_rng(int(seed))
    bw, bh = bird
    steps = rng.integers(1, max_step + 1, n_frames)
    ox = np.concatenate([[0], np.cumsum(steps)[:-1]]).astype(np.int64)
    canvas = synth_frame(h, w + int(ox[-1]) + 8, seed)
    imgs = np.empty((2 * n_frames, h, w), np.uint8)
    for i in range(n_frames):
        left = canvas[:, ox[i]:ox[i] + w]
        imgs[2 * i] = left
        imgs[2 * i + 1] = shift_frame(left, -DISPARITY, 0)
    m = max_step * n_frames + 8
    ground = synth_frame(bh + 2 * m, bw + 2 * m, seed + 1, band=False)
    pos = np.array([m, m], np.int64)
    bird_imgs = np.empty((n_frames, bh, bw), np.uint8)
    for i in range(n_frames):
        pos = np.clip(pos + rng.integers(-max_step, max_step + 1, 2), 0, 2 * m)
        bird_imgs[i] = ground[pos[1]:pos[1] + bh, pos[0]:pos[0] + bw]
    mask = np.full((bh, bw), 255, np.uint8)
    vw, vh = vehicle
    mask[bh // 2 - vh // 2:bh // 2 + vh // 2 + 1, bw // 2 - vw // 2:bw // 2 + vw // 2 + 1] = 0       # vehicle footprint (src/Frame.cc:317-327)
    f32 = np.float32
    z0 = f32(KITTI["bf"] / DISPARITY)
    poses = []
    for i in range(n_frames):
        # u_i = fx (X + tx_i) / Z0 + cx  with  u_i = u_0 - ox_i   =>  tx_i = -ox_i Z0 / fx
        t = np.array([-f32(ox[i]) * z0 / f32(KITTI["fx"]), 0, 0], f32)
        poses.append(dict(Rcw=np.eye(3, dtype=f32).reshape(9), tcw=t, Ow=(-t).astype(f32), fx=KITTI["fx"], fy=KITTI["fy"],
                          cx=float(f32(w / 2 - 13.3)), cy=float(f32(h / 2 - 2.8)), mbf=KITTI["bf"],
                          min_x=0.0, max_x=float(w), min_y=0.0, max_y=float(h),
                          log_scale_factor=float(np.log(f32(1.2), dtype=f32)), n_levels=8))
    return dict(imgs=imgs, bird_imgs=bird_imgs, bird_mask=mask, ox=ox, poses=poses, z0=float(z0))


def northstar_map(seq, extract, n_points=3000, seed=0, on_kp_frac=0.7, max_flips=40):
    """Local map of a northstar_sequence: on_kp_frac of the points are keypoints of three frames of the sequence (first,
    middle, last: `extract(img) -> (kps, desc)`) back-projected onto the scene plane through that frame's pose, with a
    perturbed copy of the keypoint's descriptor; the rest are random world points (some behind the camera, out of range,
    seen from a bad angle).  Returns dict(pos, normal, max_distance, min_distance, desc)."""
    rng = np.random.default_rng(int(seed))
    f32 = np.float32
    n = len(seq["poses"])
    h, w = seq["imgs"].shape[1:]
    z0 = seq["z0"]
    n_on = int(n_points * on_kp_frac)
    P, N, MX, D = [], [], [], []
    src_frames = sorted(set([0, n // 2, n - 1]))
    for k, fi in enumerate(src_frames):
        kps, desc = extract(seq["imgs"][2 * fi])
        take = n_on // len(src_frames) if k else n_on - (len(src_frames) - 1) * (n_on // len(src_frames))
        if len(kps) == 0:
            continue
        idx = rng.integers(0, len(kps), take)
        ps = seq["poses"][fi]
        u = kps["x"][idx].astype(np.float64) + rng.uniform(-1.5, 1.5, take)
        v = kps["y"][idx].astype(np.float64) + rng.uniform(-1.5, 1.5, take)
        Xc = np.stack([(u - ps["cx"]) * z0 / ps["fx"], (v - ps["cy"]) * z0 / ps["fy"], np.full(take, z0)], 1)
        Pw = Xc - ps["tcw"][None, :].astype(np.float64)                 # Rcw = I
        PO = Pw - ps["Ow"][None, :].astype(np.float64)
        d = np.linalg.norm(PO, axis=1)
        P.append(Pw); N.append(PO / d[:, None])
        MX.append(d * np.power(1.2, kps["octave"][idx]) * rng.uniform(0.95, 1.05, take))
        D.append(perturb_descriptors(desc[idx], max_flips, seed + 10 + k))
    n_rnd = n_points - sum(len(p) for p in P)
    if n_rnd > 0:
        z = rng.uniform(5.0, 120.0, n_rnd) * np.where(rng.random(n_rnd) < 0.15, -1.0, 1.0)
        ps = seq["poses"][n // 2]
        u = rng.uniform(-0.3 * w, 1.3 * w, n_rnd)
        v = rng.uniform(-0.3 * h, 1.3 * h, n_rnd)
        Xc = np.stack([(u - ps["cx"]) * z / ps["fx"], (v - ps["cy"]) * z / ps["fy"], z], 1)
        Pw = Xc - ps["tcw"][None, :].astype(np.float64)
        PO = Pw - ps["Ow"][None, :].astype(np.float64)
        d = np.linalg.norm(PO, axis=1)
        nrm = PO / np.maximum(d, 1e-9)[:, None] + rng.normal(0, 0.5, (n_rnd, 3)) * (rng.random(n_rnd) < 0.3)[:, None] * 3.0
        P.append(Pw); N.append(nrm / np.linalg.norm(nrm, axis=1)[:, None])
        MX.append(np.abs(d) * np.power(1.2, rng.integers(0, 8, n_rnd)) * np.where(rng.random(n_rnd) < 0.1, rng.choice([0.2, 6.0], n_rnd), 1.0))
        D.append(synth_descriptors(n_rnd, seed + 20))
    pos = np.concatenate(P).astype(f32)
    maxd = np.concatenate(MX).astype(f32)
    perm = rng.permutation(len(pos))
    return dict(pos=np.ascontiguousarray(pos[perm]), normal=np.ascontiguousarray(np.concatenate(N).astype(f32)[perm]),
                max_distance=np.ascontiguousarray(maxd[perm]), min_distance=np.ascontiguousarray((maxd / f32(1.2 ** 7)).astype(f32)[perm]),
                desc=np.ascontiguousarray(np.concatenate(D).astype(np.uint8)[perm]))
