"""bench.py contract checks that need no GPU: the reference arm (the reference's own ORBextractor.cc from oracle/_ref + the oracle
port of the matcher, on the host cores) prints ONE JSON line with the keys the driver reads, the same `config` / `metric` / `unit`
as the CUDA arm, and zero-byte `e2e` copies."""
import json
import os
import subprocess
import sys

from helpers import ROOT


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-frames-per-step", "4"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["metric"] == "ORB extract+match frames/sec" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["dtype"] == "u8" and d["data"] == "synthetic"
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"]
    # the CUDA arm prints the same config (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_dict(128, 3)
    assert d["gpu_launches"] == 0


def test_rank_placement_is_a_distinct_gpu_per_rank(monkeypatch):
    """bench.place_rank: fewer ranks than GPUs are spread over the two halves of the box (0, 4, 1, 5, ...), every rank gets its own GPU,
    as many ranks as GPUs (or ORBB200_RANK_PLACEMENT=packed) is the identity."""
    import bench
    monkeypatch.delenv("ORBB200_RANK_PLACEMENT", raising=False)
    for ndev in (2, 4, 8, 16):
        for world in range(1, ndev + 1):
            gpus = [bench.place_rank(l, world, ndev)[0] for l in range(world)]
            assert len(set(gpus)) == world and all(0 <= g < ndev for g in gpus), (ndev, world, gpus)
            if world == ndev or world == 1 or ndev < 4:
                assert gpus == list(range(world))
    assert [bench.place_rank(l, 4, 8)[0] for l in range(4)] == [0, 4, 1, 5]
    assert [bench.place_rank(l, 2, 8)[0] for l in range(2)] == [0, 4]
    assert bench.place_rank(3, 4, 7) == (3, "packed")                # odd GPU counts: no halves
    monkeypatch.setenv("ORBB200_RANK_PLACEMENT", "packed")
    assert [bench.place_rank(l, 4, 8)[0] for l in range(4)] == [0, 1, 2, 3]
