"""Import the UNMODIFIED reference (fatiando/verde) from /root/reference.

TEST INFRASTRUCTURE ONLY. Works only in the build container, where
/root/reference exists; it does not exist on the GPU box. Used by
oracle/make_golden.py (fixture generation) and by CPU tests that pin the
oracle restatement against the live reference (skipped when absent).

The reference imports four things at import time that this image lacks
(SURVEY.md section 8c): ``verde._version_generated`` (setuptools_scm output),
``pooch``, ``dask`` and ``xarray``. They are replaced by inert in-memory
modules; none of them is on the arithmetic path of Spline / VectorSpline2D /
cross_val_score / SplineCV (non-delayed).
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "verde"))


def import_reference():
    """Return the reference ``verde`` module (raises ImportError if absent)."""
    if not reference_available():
        raise ImportError("reference checkout not present at " + REFERENCE_ROOT)
    if "verde" in sys.modules and getattr(sys.modules["verde"], "__file__", "").startswith(REFERENCE_ROOT):
        return sys.modules["verde"]
    for name in ("dask", "xarray", "pooch"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    dask = sys.modules["dask"]
    if not hasattr(dask, "delayed"):
        dask.delayed = lambda f: f
    pooch = sys.modules["pooch"]
    if not hasattr(pooch, "create"):
        pooch.os_cache = lambda *a, **k: "/tmp/verde-pooch"
        pooch.create = lambda *a, **k: None
    gen = types.ModuleType("verde._version_generated")
    gen.version = "0+reference"
    sys.modules["verde._version_generated"] = gen
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import verde  # noqa: E402

    return verde
