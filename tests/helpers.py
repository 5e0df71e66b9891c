"""Shared imports for the tests: the oracle (checker), the synthetic generator and the product package."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle  # noqa: E402  (test infrastructure)
from oracle import KP_DTYPE  # noqa: E402,F401


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


synth = sys.modules.get("orbb200_synth") or _load("orbb200_synth", os.path.join(ROOT, "orb-slam-birdview_b200", "synth.py"))
GOLDEN = os.path.join(ROOT, "tests", "golden")
