"""How often does the documented octree tie rule equal what the REFERENCE BINARY does?  

Runs oracle/_ref (the reference's src/ORBextractor.cc compiled unmodified, three builds: see tests/test_oracle_ref.py)
next to the oracle on synthetic frames of the config shapes and writes profiles/r2_ref_tie_report.json.
Test infrastructure: CPU only, nothing here is on the product path.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from importlib import import_module  # noqa: E402

synth = import_module("orb_slam_birdview_b200.synth")

SHAPES = {"C1_752x480_1000": (480, 752, 1000, 20, 7), "C2_1241x376_2000": (376, 1241, 2000, 20, 7),
          "C3bird_400x400_2000": (400, 400, 2000, 15, 5), "C5_1920x1080_4000": (1080, 1920, 4000, 20, 7)}


def keyset(k):
    return set(zip(k["x"].tolist(), k["y"].tolist(), k["octave"].tolist()))


def main(frames=8):
    out = {}
    for name, (h, w, nf, ini, mn) in SHAPES.items():
        O = oracle.Extractor(nf, 1.2, 8, ini, mn)
        Rb = oracle.RefExtractor(nf, 1.2, 8, ini, mn, variant="bump")
        Rg = oracle.RefExtractor(nf, 1.2, 8, ini, mn, variant="glibc")
        Rn = oracle.RefExtractor(nf, 1.2, 8, ini, mn, variant="nofma")
        r = dict(frames=frames, bump_identical_frames=0, levels=0, levels_identical_glibc=0, levels_same_set_glibc=0,
                 kps=0, kps_shared_glibc=0, kps_shared_between_two_glibc_runs=0, glibc_rerun_identical_frames=0,
                 shared_kps_record_mismatch=0, shared_desc_mismatch_nofma=0, shared_desc_mismatch_fma=0, shared_desc=0)
        for f in range(frames):
            img = synth.synth_frame(h, w, 9000 + 17 * f)
            k0, d0 = O(img)
            kb, db = Rb(img)
            r["bump_identical_frames"] += int(kb.tobytes() == k0.tobytes() and np.array_equal(db, d0))
            kg, dg = Rg(img)
            kg2, _ = Rg(img)
            kn, dn = Rn(img)
            r["glibc_rerun_identical_frames"] += int(kg.tobytes() == kg2.tobytes())
            r["kps"] += len(k0)
            r["kps_shared_glibc"] += len(keyset(kg) & keyset(k0))
            r["kps_shared_between_two_glibc_runs"] += len(keyset(kg) & keyset(kg2))
            for lvl in range(8):
                a, b = kg[kg["octave"] == lvl], k0[k0["octave"] == lvl]
                r["levels"] += 1
                r["levels_identical_glibc"] += int(a.tobytes() == b.tobytes())
                r["levels_same_set_glibc"] += int(keyset(a) == keyset(b))
            idx0 = {key: i for i, key in enumerate(zip(k0["x"].tolist(), k0["y"].tolist(), k0["octave"].tolist()))}
            for (k, d, tag) in ((kn, dn, "nofma"), (kg, dg, "fma")):
                for i, key in enumerate(zip(k["x"].tolist(), k["y"].tolist(), k["octave"].tolist())):
                    j = idx0.get(key)
                    if j is None:
                        continue
                    if tag == "fma":
                        r["shared_desc"] += 1
                        r["shared_kps_record_mismatch"] += int(k[i].tobytes() != k0[j].tobytes())
                    r[f"shared_desc_mismatch_{tag}"] += int(not np.array_equal(d[i], d0[j]))
        r["tie_rule_level_agreement"] = r["levels_identical_glibc"] / r["levels"]
        r["keypoint_set_agreement"] = r["kps_shared_glibc"] / r["kps"]
        r["reference_self_agreement_between_two_runs"] = r["kps_shared_between_two_glibc_runs"] / r["kps"]
        r["fma_descriptor_mismatch_rate"] = r["shared_desc_mismatch_fma"] / max(r["shared_desc"], 1)
        out[name] = r
        print(name, json.dumps(r))
    doc = {"what": "oracle (= CUDA path) vs the reference's own ORBextractor.cc compiled unmodified (oracle/_ref)",
           "builds": {"bump": "monotone operator new, -ffp-contract=off", "glibc": "glibc malloc, FMA contraction on (-O3 -march=x86-64-v3)",
                      "nofma": "glibc malloc, -ffp-contract=off"},
           "shapes": out}
    with open(os.path.join(ROOT, "profiles", "r2_ref_tie_report.json"), "w") as f:
        json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
