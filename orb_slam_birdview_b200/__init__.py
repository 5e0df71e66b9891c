"""Import alias: the package directory is `orb-slam-birdview_b200/` (not a Python identifier)."""
import os

__path__ = [os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orb-slam-birdview_b200")]
with open(os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), os.path.join(__path__[0], "__init__.py"), "exec"))
