"""A small synthetic SLAM scene for the ORBmatcher suite drivers (orb-slam-birdview_b200/cpp/matcher_suite.cpp): map points in
front of four views (last frame, current frame, two keyframes), keypoints at their (noisy) projections with perturbed copies of
the points' descriptors plus clutter, DBoW2-style feature vectors, a birdview layer with ground landmarks, and the arguments of
every ORBmatcher call.  `write_scene` serialises it in the order `read_scene()` of the driver reads; `parse_output` reads what
the driver wrote.  numpy only, seeded."""
import struct

import numpy as np

from helpers import KP_DTYPE, synth

W, H = 1241, 376
FX, FY, CX, CY, BF = 718.856, 718.856, 607.1928, 185.2157, 386.1448
NLEVELS, SCALE = 8, np.float32(1.2)
BIRD = 400
F32 = np.float32


def _rot(ax, ay, az):
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _pose(rng, rot=0.02, trans=0.4):
    T = np.eye(4)
    T[:3, :3] = _rot(*rng.normal(0, rot, 3))
    T[:3, 3] = rng.normal(0, trans, 3)
    return T.astype(F32)


def _featvec(nodes):
    """CSR over ascending node ids: (node[nn], ptr[nn+1], idx[...]) -- DBoW2::FeatureVector iteration order"""
    order = np.argsort(nodes, kind="stable")
    ids, counts = np.unique(nodes, return_counts=True)
    return ids.astype(np.int32), np.concatenate([[0], np.cumsum(counts)]).astype(np.int32), order.astype(np.int32)


def make_scene(seed=7, n_points=900, n_clutter=250, n_birds=500, kf_assoc=0.85, only_stereo=0, params=None):
    """params: optional overrides of the 14 float call parameters by position (see `args['params']` below)"""
    rng = np.random.default_rng(seed)
    sf = np.array([F32(1.2) ** 0] * NLEVELS, F32)
    s = F32(1)
    for i in range(NLEVELS):
        sf[i] = s
        s = F32(s * SCALE)
    # ---- world points, in front of an identity camera ----
    z = rng.uniform(5, 60, n_points)
    u = rng.uniform(-50, W + 50, n_points)
    v = rng.uniform(-30, H + 30, n_points)
    Pw = np.stack([(u - CX) * z / FX, (v - CY) * z / FY, z], 1).astype(F32)
    base_desc = synth.synth_descriptors(n_points, seed + 1)
    base_angle = rng.uniform(0, 360, n_points)
    node_of = (np.arange(n_points) % 50).astype(np.int32)
    ref_level = rng.integers(0, 6, n_points)
    dist0 = np.linalg.norm(Pw, axis=1)
    maxd = (dist0 * np.power(1.2, ref_level) * rng.uniform(0.95, 1.05, n_points)).astype(F32)
    poses = [_pose(rng) for _ in range(4)]            # last frame, current frame, keyframe 0, keyframe 1
    views = []
    for k, T in enumerate(poses):
        Pc = (Pw.astype(np.float64) @ T[:3, :3].T.astype(np.float64)) + T[:3, 3].astype(np.float64)
        zz = Pc[:, 2]
        uu = FX * Pc[:, 0] / zz + CX
        vv = FY * Pc[:, 1] / zz + CY
        vis = (zz > 1.0) & (uu > 20) & (uu < W - 20) & (vv > 20) & (vv < H - 20) & (rng.random(n_points) < 0.8)
        ids = np.nonzero(vis)[0]
        n_true = len(ids)
        N = n_true + n_clutter
        kps = np.zeros(N, KP_DTYPE)
        Ow = -(T[:3, :3].T.astype(np.float64) @ T[:3, 3].astype(np.float64))
        d = np.linalg.norm(Pw[ids].astype(np.float64) - Ow, axis=1)
        pred = np.clip(np.ceil(np.log(maxd[ids] / d) / np.log(1.2)), 0, NLEVELS - 1).astype(np.int64)
        octv = np.clip(pred - rng.integers(0, 2, n_true) - (rng.random(n_true) < 0.05) * 2, 0, NLEVELS - 1)
        kps["x"][:n_true] = (uu[ids] + rng.normal(0, 0.7, n_true)).astype(F32)
        kps["y"][:n_true] = (vv[ids] + rng.normal(0, 0.7, n_true)).astype(F32)
        kps["octave"][:n_true] = octv
        kps["angle"][:n_true] = ((base_angle[ids] + 17.0 * k + rng.normal(0, 3, n_true)) % 360).astype(F32)
        kps["x"][n_true:] = rng.uniform(20, W - 20, n_clutter).astype(F32)
        kps["y"][n_true:] = rng.uniform(20, H - 20, n_clutter).astype(F32)
        kps["octave"][n_true:] = rng.integers(0, NLEVELS, n_clutter)
        kps["angle"][n_true:] = rng.uniform(0, 360, n_clutter).astype(F32)
        kps["size"] = (31 * sf[kps["octave"]]).astype(np.int32).astype(F32)
        kps["response"] = rng.integers(7, 120, N).astype(F32)
        kps["class_id"] = -1
        desc = np.concatenate([synth.perturb_descriptors(base_desc[ids], 30, seed + 10 + k), synth.synth_descriptors(n_clutter, seed + 20 + k)])
        ur = np.full(N, -1, F32)
        st = rng.random(n_true) < 0.6
        ur[:n_true][st] = (kps["x"][:n_true][st] - BF / zz[ids][st] + rng.normal(0, 0.3, int(st.sum()))).astype(F32)
        mp = np.full(N, -1, np.int32)
        assoc = {0: 0.7, 1: 0.1, 2: kf_assoc, 3: kf_assoc}[k]
        has = rng.random(n_true) < assoc
        mp[:n_true][has] = ids[has]
        nodes = np.concatenate([node_of[ids], rng.integers(0, 60, n_clutter)]).astype(np.int32)
        outl = np.zeros(N, np.uint8)
        outl[:n_true] = rng.random(n_true) < 0.05
        perm = rng.permutation(N)
        views.append(dict(id=100 + k, T=T, kps=np.ascontiguousarray(kps[perm]), desc=np.ascontiguousarray(desc[perm]), ur=np.ascontiguousarray(ur[perm]),
                          mp=np.ascontiguousarray(mp[perm]), outl=np.ascontiguousarray(outl[perm]), fv=_featvec(nodes[perm]), Ow=Ow))
    # ---- map point records ----
    Ow2 = views[2]["Ow"]
    PO = Pw.astype(np.float64) - Ow2
    normal = (PO / np.linalg.norm(PO, axis=1)[:, None]).astype(F32)
    Tc = poses[1]
    Pc = (Pw.astype(np.float64) @ Tc[:3, :3].T.astype(np.float64)) + Tc[:3, 3].astype(np.float64)
    pu = (FX * Pc[:, 0] / Pc[:, 2] + CX).astype(F32)
    pv = (FY * Pc[:, 1] / Pc[:, 2] + CY).astype(F32)
    inview = (Pc[:, 2] > 0) & (pu > 0) & (pu < W) & (pv > 0) & (pv < H) & (rng.random(n_points) < 0.9)
    dcur = np.linalg.norm(Pw.astype(np.float64) + (Tc[:3, :3].T.astype(np.float64) @ Tc[:3, 3].astype(np.float64)), axis=1)
    mps = dict(pos=Pw, normal=normal, mind=(maxd / F32(1.2 ** 7)).astype(F32), maxd=maxd, desc=synth.perturb_descriptors(base_desc, 10, seed + 2),
               nobs=rng.integers(0, 6, n_points).astype(np.int32), bad=(rng.random(n_points) < 0.03).astype(np.int32), inview=inview.astype(np.int32),
               projx=pu, projy=pv, projxr=(pu - BF / Pc[:, 2]).astype(F32), viewcos=np.where(rng.random(n_points) < 0.5, 0.999, 0.9).astype(F32),
               level=np.clip(np.ceil(np.log(maxd / dcur) / np.log(1.2)), 0, NLEVELS - 1).astype(np.int32),
               lastseen=np.where(rng.random(n_points) < 0.1, 101, 0).astype(np.int32))
    # ---- birdview: landmarks on the ground in the body frame of the current frame ----
    Tbc = np.eye(4, dtype=F32)
    Tbc[:3, 3] = [0.1, -0.05, 0.02]
    Tbw = Tbc.astype(np.float64) @ poses[1].astype(np.float64)
    local = np.stack([rng.uniform(-5.5, 8.5, n_birds), rng.uniform(-7, 7, n_birds), rng.normal(0, 0.08, n_birds), np.ones(n_birds)], 1)
    world = (np.linalg.inv(Tbw) @ local.T).T[:, :3].astype(F32)
    bdesc = synth.synth_descriptors(n_birds, seed + 30)
    px = BIRD / 2 - local[:, 1] * 25.1
    py = BIRD / 2 - (local[:, 0] - 1.393) * 25.1
    birds = dict(pos=world, desc=synth.perturb_descriptors(bdesc, 8, seed + 31), nobs=rng.integers(0, 4, n_birds).astype(np.int32),
                 lastseen=np.where(rng.random(n_birds) < 0.1, 101, 0).astype(np.int32))
    boct = rng.integers(0, 4, n_birds)
    bang = rng.uniform(0, 360, n_birds)
    for k, view in enumerate(views):
        shift = {0: (3.0, -2.0), 1: (0.0, 0.0), 2: (-4.0, 5.0), 3: (0, 0)}[k]
        vis = (px > 5) & (px < BIRD - 5) & (py > 5) & (py < BIRD - 5) & (rng.random(n_birds) < 0.8)
        ids = np.nonzero(vis)[0]
        nt, nc = len(ids), 120
        bk = np.zeros(nt + nc, KP_DTYPE)
        bk["x"][:nt] = (px[ids] + shift[0] + rng.normal(0, 0.6, nt)).astype(F32)
        bk["y"][:nt] = (py[ids] + shift[1] + rng.normal(0, 0.6, nt)).astype(F32)
        bk["octave"][:nt] = boct[ids]
        bk["angle"][:nt] = ((bang[ids] + 11.0 * k + rng.normal(0, 3, nt)) % 360).astype(F32)
        bk["x"][nt:] = rng.uniform(5, BIRD - 5, nc).astype(F32)
        bk["y"][nt:] = rng.uniform(5, BIRD - 5, nc).astype(F32)
        bk["octave"][nt:] = rng.integers(0, 4, nc)
        bk["angle"][nt:] = rng.uniform(0, 360, nc).astype(F32)
        bk["size"], bk["response"], bk["class_id"] = 31, 20, -1
        bd = np.concatenate([synth.perturb_descriptors(bdesc[ids], 25, seed + 40 + k), synth.synth_descriptors(nc, seed + 50 + k)])
        bmp = np.full(nt + nc, -1, np.int32)
        has = rng.random(nt) < {0: 0.7, 1: 0.1, 2: 0.8, 3: 0.0}[k]
        bmp[:nt][has] = ids[has]
        perm = rng.permutation(nt + nc)
        view.update(bk=np.ascontiguousarray(bk[perm]), bdesc=np.ascontiguousarray(bd[perm]), bmp=np.ascontiguousarray(bmp[perm]))
    # ---- call arguments ----
    kf0, kf1 = views[2], views[3]
    s_loop = 1.1
    Scw = poses[2].astype(np.float64).copy()
    Scw[:3, :] *= s_loop
    T12 = poses[2].astype(np.float64) @ np.linalg.inv(poses[3].astype(np.float64))
    R12, t12 = T12[:3, :3], T12[:3, 3]
    K = np.array([[FX, 0, CX], [0, FY, CY], [0, 0, 1]])
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    F12 = np.linalg.inv(K).T @ tx @ R12 @ np.linalg.inv(K)                     # LocalMapping::ComputeF12
    in_kf0 = kf0["mp"][kf0["mp"] >= 0]
    in_kf1 = set(kf1["mp"][kf1["mp"] >= 0].tolist())
    scw_matched = np.where(rng.random(len(kf0["mp"])) < 0.2, kf0["mp"], -1).astype(np.int32)
    sim3 = np.full(len(kf0["mp"]), -1, np.int32)
    for i, m in enumerate(kf0["mp"]):
        if m >= 0 and m in in_kf1 and rng.random() < 0.15:
            sim3[i] = m
    fuse = rng.integers(0, n_points, 700).astype(np.int32)
    fuse[rng.random(700) < 0.05] = -1
    fuse[100:110] = fuse[0:10]                                                  # the same MapPoint twice in the list
    f0 = views[0]
    o0 = np.nonzero(f0["kps"]["octave"] == 0)[0]
    init_prev = np.stack([f0["kps"]["x"], f0["kps"]["y"]], 1).astype(F32)
    bird_prev = np.stack([f0["bk"]["x"] + F32(1.5), f0["bk"]["y"] - F32(1.0)], 1).astype(F32)
    args = dict(params=np.array([3.0, 0.8, 15.0, 10.0, 10.0, 3.0, 7.5, 4.0, 15.0, 0.7, 0.9, 0.99, 0.6, 1.0], F32),
                ints=np.array([100, 100, 15, only_stereo], np.int32), Scw=Scw.astype(F32), R12=R12.astype(F32), t12=t12.astype(F32), F12=F12.astype(F32),
                already_found=in_kf0[:30].astype(np.int32), scw_points=rng.integers(0, n_points, 500).astype(np.int32), scw_matched=scw_matched,
                fuse_points=fuse, fuse_scw_points=rng.integers(0, n_points, 500).astype(np.int32), sim3_matches12=sim3,
                bird_proj_points=np.arange(n_birds, dtype=np.int32), init_prev=init_prev, bird_prev=bird_prev, octave0_count=len(o0))
    for k, v in (params or {}).items():
        args["params"][k] = v
    return dict(sf=sf, views=views, mps=mps, birds=birds, Tbc=Tbc, args=args)


def write_scene(S, path):
    out = bytearray()

    def i32(*v):
        out.extend(struct.pack(f"<{len(v)}i", *[int(x) for x in v]))

    def f32(a):
        out.extend(np.ascontiguousarray(a, F32).tobytes())

    def raw(a):
        out.extend(np.ascontiguousarray(a).tobytes())

    i32(NLEVELS)
    f32([SCALE])
    f32([0.0, W, 0.0, H, 64.0 / W, 48.0 / H, 64.0 / BIRD, 48.0 / BIRD])
    i32(BIRD, BIRD)
    f32(S["Tbc"])
    m = S["mps"]
    n = len(m["pos"])
    i32(n)
    for i in range(n):
        f32(m["pos"][i]); f32(m["normal"][i]); f32([m["mind"][i], m["maxd"][i]]); raw(m["desc"][i])
        i32(m["nobs"][i], m["bad"][i], m["inview"][i])
        f32([m["projx"][i], m["projy"][i], m["projxr"][i], m["viewcos"][i]])
        i32(m["level"][i], m["lastseen"][i])
    b = S["birds"]
    i32(len(b["pos"]))
    for i in range(len(b["pos"])):
        f32(b["pos"][i]); raw(b["desc"][i]); i32(b["nobs"][i], b["lastseen"][i])
    for v in S["views"]:
        i32(v["id"], len(v["kps"]), len(v["bk"]))
        f32([FX, FY, CX, CY, BF, BF / FX])
        f32(v["T"])
        raw(v["kps"]); raw(v["desc"]); f32(v["ur"]); raw(v["mp"].astype(np.int32)); raw(v["outl"].astype(np.uint8))
        node, ptr, idx = v["fv"]
        i32(len(node)); raw(node); raw(ptr); raw(idx)
        raw(v["bk"]); raw(v["bdesc"]); raw(v["bmp"].astype(np.int32))
    a = S["args"]
    f32(a["params"]); raw(a["ints"]); f32(a["Scw"]); f32(a["R12"]); f32(a["t12"]); f32(a["F12"])
    for k in ("already_found", "scw_points", "scw_matched", "fuse_points", "fuse_scw_points", "sim3_matches12", "bird_proj_points"):
        i32(len(a[k])); raw(a[k].astype(np.int32))
    for k in ("init_prev", "bird_prev"):
        i32(len(a[k])); f32(a[k])
    with open(path, "wb") as f:
        f.write(out)


def parse_output(path):
    """-> dict tag -> dict(ret=..., arrays...)"""
    data = open(path, "rb").read()
    pos = 0

    def tag():
        nonlocal pos
        t = data[pos:pos + 8].rstrip(b"\0").decode()
        pos += 8
        return t

    def i():
        nonlocal pos
        v = struct.unpack_from("<i", data, pos)[0]
        pos += 4
        return v

    def ints():
        nonlocal pos
        n = i()
        a = np.frombuffer(data, "<i4", n, pos).copy()
        pos += 4 * n
        return a

    def pts():
        nonlocal pos
        n = i()
        a = np.frombuffer(data, "<f4", 2 * n, pos).copy().reshape(-1, 2)
        pos += 8 * n
        return a

    def log():
        n = i()
        return np.array([[i(), i(), i(), i()] for _ in range(n)], np.int32).reshape(-1, 4)

    out = {}
    while pos < len(data):
        t = tag()
        r = dict(ret=i())
        if t in ("SBP_MPS", "SBP_FFS", "SBP_FFM", "SBP_FKF", "SBP_SCW", "BOW_KFF", "BOW_KK", "SIM3", "BIRD_PRJ", "BIRD_FF", "BIRD_KF"):
            r["out"] = ints()
        elif t in ("INIT", "BIRD_PRV"):
            r["out"] = ints(); r["prev"] = pts()
        elif t == "BIRD":
            r["out"] = ints()
        elif t == "TRIANG":
            n = i()
            r["pairs"] = np.array([[i(), i()] for _ in range(n)], np.int32).reshape(-1, 2)
        elif t == "FUSE":
            r["kf"] = ints(); r["log"] = log()
        elif t == "FUSE_SCW":
            r["repl"] = ints(); r["kf"] = ints(); r["log"] = log()
        elif t == "DIST":
            pass
        else:
            raise ValueError(f"unknown tag {t!r} at {pos}")
        out[t] = r
    return out
