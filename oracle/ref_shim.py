"""TEST INFRASTRUCTURE ONLY -- import shim for the *Python reference* (amansanghvi/Thesis).

Only usable where /root/reference exists (the build container).  It is used by
tests/golden/make_golden.py to generate the committed golden vectors and by the
`not gpu` tests that pin oracle/rbpf_oracle.c against the reference's own code.
Nothing in thesis_b200/ may import this module.

The reference modules import `matplotlib.pyplot` and `matlab` at top level
(gridmap.py:4-6, hybridmap.py:5-7, robot.py:3, lidar.py:5, main.py:4-5); both
are absent here, so empty stand-ins are injected into sys.modules first.
Loaders open "./data/..." relative paths (IntelLidarData.py:13), so callers
that construct loaders must run them under `ref_cwd()`.
"""
import contextlib
import os
import sys
import types

REF_ROOT = os.environ.get("THESIS_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "hybridmap.py"))


def _stub(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def install():
    """Make `import robot, hybridmap, gridmap, lidar, models, main` work."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if "matlab" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        ml = _stub("matlab")
        ml.double = lambda x: x          # hybridmap.py:245 wraps lists in matlab.double
        ml.engine = _stub("matlab.engine")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


@contextlib.contextmanager
def ref_cwd():
    old = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


def fresh_hybridmap(matlab_obj="stub"):
    """A HybridMap with a *private* tile list (the class-level
    `_maps` list at hybridmap.py:64 is shared by every instance; we reset it so
    each particle owns its map, a declared deviation)."""
    install()
    import hybridmap
    hybridmap.HybridMap._maps = []
    m = hybridmap.HybridMap(matlab_obj)
    m._maps = list(m._maps)              # instance attribute, detached from the class list
    hybridmap.HybridMap._maps = []
    return m


def fresh_robot(matlab_obj="stub"):
    install()
    import robot
    r = robot.Robot(None)
    import numpy as np
    r._map = fresh_hybridmap(matlab_obj)
    r._weight = [1.0]
    r._cov = np.zeros((3, 3), dtype=np.float64)   # pinned to float64
    r._x = [0.0]
    r._y = [0.0]
    r._theta = [0.0]
    return r
