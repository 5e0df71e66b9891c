"""Every ORB_SLAM2::ORBmatcher method, called through Frame / KeyFrame / MapPoint / MapPointBird objects the way Tracking,
LocalMapping and LoopClosing call them (orb-slam-birdview_b200/cpp/matcher_suite.cpp), three ways on the same scene files:

  oracle/_ref/matcher_suite_ref      the reference's own, UNMODIFIED src/ORBmatcher.cc (compiled where it lies by
                                     oracle/ref_standin/Makefile against the compat object model)      -> the expected output
  oracle/_ref/matcher_suite_oracle   the product's adapters (cpp/ORBmatcher_b200.cc) over the C ABI implemented on the CPU
                                     oracle (oracle/abi_on_oracle.cpp): pins the oracle's matcher loops and the adapters' host
                                     geometry on the real reference code, without a GPU                 (CPU test)
  cpp/matcher_suite                  the same adapters over liborbb200.so                               (GPU test: the product)

Compared: return values and everything the calls leave behind (mvpMapPoints, vpMatched, vnMatches12, vbPrevMatched, matched
pairs, vpReplacePoint, and the log of AddObservation / AddMapPoint / Replace calls of Fuse), all exact.
tests/golden/matcher_suite_ref.npz holds the reference binary's output for the default scene (written by this file's
`python tests/test_matcher_suite.py --write-golden` in the container that has /root/reference)."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

import matcher_scene as ms
from helpers import ROOT

REF = os.path.join(ROOT, "oracle", "_ref", "matcher_suite_ref")
ORA = os.path.join(ROOT, "oracle", "_ref", "matcher_suite_oracle")
GPU = os.path.join(ROOT, "orb-slam-birdview_b200", "cpp", "matcher_suite")
GOLDEN = os.path.join(ROOT, "tests", "golden", "matcher_suite_ref.npz")

SCENES = {
    "default": dict(seed=7),
    "sparse_kf": dict(seed=11, n_points=1400, n_clutter=400, n_birds=800, kf_assoc=0.5),           # many unassociated keypoints: SearchForTriangulation
    "only_stereo": dict(seed=12, n_points=700, n_clutter=100, n_birds=300, kf_assoc=0.6, only_stereo=1),
    "wide": dict(seed=13, n_points=1200, n_clutter=600, n_birds=600,                               # wide windows, loose ratios: long greedy chains
                 params={0: 6.0, 1: 0.95, 2: 30.0, 3: 20.0, 4: 20.0, 5: 6.0, 6: 12.0, 7: 10.0, 8: 25.0, 9: 0.9, 10: 0.95, 12: 0.9}),
    "tight": dict(seed=14, n_points=500, n_clutter=50, n_birds=200, params={0: 1.0, 1: 0.6, 2: 7.0, 5: 2.0, 9: 0.6}),
    "tiny": dict(seed=15, n_points=40, n_clutter=6, n_birds=24),                                   # a few dozen features: most calls find little or nothing
}
SPARSE = {"tiny"}                                                                                  # scenes whose calls may legitimately come back empty


def _have_ref_build():
    if not (os.path.exists(REF) and os.path.exists(ORA)) and os.path.exists("/root/reference/src/ORBmatcher.cc"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle", "ref_standin")], check=True)
    return os.path.exists(REF) and os.path.exists(ORA)


def _run(binary, scene_path, out_path):
    subprocess.run([binary, str(scene_path), str(out_path)], check=True, timeout=300)
    return ms.parse_output(out_path)


def _same(a, b, name):
    assert list(a) == list(b), name
    for tag in a:
        assert a[tag]["ret"] == b[tag]["ret"], f"{name}: {tag} returned {b[tag]['ret']}, reference {a[tag]['ret']}"
        for f in a[tag]:
            if f != "ret":
                assert np.array_equal(a[tag][f], b[tag][f]), f"{name}: {tag}.{f} differs from the reference"


def _flatten(out):
    return {f"{tag}.{f}": np.asarray(v) for tag, r in out.items() for f, v in r.items()}


@pytest.mark.parametrize("name", list(SCENES))
def test_adapters_on_oracle_equal_unmodified_reference(name, tmp_path):
    """CPU: adapters + oracle == the reference's own ORBmatcher.cc, for all 16 methods + DescriptorDistance."""
    if not _have_ref_build():
        pytest.skip("oracle/_ref matcher binaries not built (needs /root/reference)")
    scene = tmp_path / "scene.bin"
    ms.write_scene(ms.make_scene(**SCENES[name]), scene)
    ref = _run(REF, scene, tmp_path / "ref.bin")
    ora = _run(ORA, scene, tmp_path / "ora.bin")
    _same(ref, ora, name)
    # the scene exercises every method: no call may come back empty
    for tag, r in ref.items():
        assert r["ret"] > 0 or name in SPARSE, f"{name}: {tag} found nothing -- the scene does not exercise it"


def test_reference_output_equals_committed_golden(tmp_path):
    """The committed golden output of the default scene is what the reference binary produces (when it is present)."""
    g = np.load(GOLDEN)
    scene = tmp_path / "scene.bin"
    ms.write_scene(ms.make_scene(**SCENES["default"]), scene)
    if hashlib.sha1(open(scene, "rb").read()).hexdigest() != str(g["scene_sha1"]):
        pytest.skip("numpy on this host generates a different default scene than the one the golden file was made from")
    if not _have_ref_build():
        pytest.skip("oracle/_ref matcher binaries not built (needs /root/reference)")
    flat = _flatten(_run(REF, scene, tmp_path / "ref.bin"))
    for k, v in flat.items():
        assert np.array_equal(g[k], v), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SCENES))
def test_product_adapters_equal_unmodified_reference(name, tmp_path):
    """GPU: cpp/ORBmatcher_b200.cc over liborbb200.so == the reference's own ORBmatcher.cc (prebuilt oracle/_ref binary; for the
    default scene the committed golden output when the binary did not travel)."""
    if not os.path.exists(GPU):
        subprocess.run(["make", "-s", "-C", os.path.dirname(GPU), "matcher_suite"], check=True)
    scene = tmp_path / "scene.bin"
    ms.write_scene(ms.make_scene(**SCENES[name]), scene)
    got = _run(GPU, scene, tmp_path / "gpu.bin")
    if os.path.exists(REF):
        _same(_run(REF, scene, tmp_path / "ref.bin"), got, name)
        return
    g = np.load(GOLDEN)
    if name != "default" or hashlib.sha1(open(scene, "rb").read()).hexdigest() != str(g["scene_sha1"]):
        pytest.skip("no reference binary on this box and no golden output for this scene")
    for k, v in _flatten(got).items():
        assert np.array_equal(g[k], v), k


if __name__ == "__main__" and "--write-golden" in sys.argv:
    import tempfile
    assert _have_ref_build()
    with tempfile.TemporaryDirectory() as d:
        scene = os.path.join(d, "scene.bin")
        ms.write_scene(ms.make_scene(**SCENES["default"]), scene)
        flat = _flatten(_run(REF, scene, os.path.join(d, "ref.bin")))
        np.savez_compressed(GOLDEN, scene_sha1=hashlib.sha1(open(scene, "rb").read()).hexdigest(), **flat)
        print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")
