import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import oracle, synth
import orb_slam_birdview_b200 as pkg

h, w, nf = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (480, 752, 1000)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
img = synth.synth_frame(h, w, 1000)
ex = pkg.ORBextractor(nf, 1.2, 8, 20, 7, max_size=(w, h))
orc = oracle.Extractor(nf, 1.2, 8, 20, 7)
k0, d0 = orc(img)
for rep in range(reps):
    k, d = ex(img)
    msg = []
    for lvl in range(8):
        c = ex.level_candidates(0, lvl); c0 = orc.level_candidates(lvl)
        key = lambda c: np.lexsort((c[:, 2], c[:, 0], c[:, 1]))
        if len(c) != len(c0) or not np.array_equal(c[key(c)], c0[key(c0)]):
            msg.append(f"cand L{lvl} {len(c)} vs {len(c0)}")
    if len(k) != len(k0):
        msg.append(f"count {len(k)} vs {len(k0)}")
    else:
        for f in k.dtype.names:
            bad = np.nonzero(k[f] != k0[f])[0]
            if len(bad):
                msg.append(f"{f}: {len(bad)} bad, first {bad[:4]} oct {k0['octave'][bad[:4]]} got {k[f][bad[:4]]} want {k0[f][bad[:4]]}")
        bd = np.nonzero((d != d0).any(1))[0]
        if len(bd):
            msg.append(f"desc rows bad {len(bd)} first {bd[:4]}")
    print("rep", rep, "OK" if not msg else msg)
