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
