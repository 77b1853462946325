import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def record_parity(name, payload):
    """Append the observed GPU-vs-reference gaps of a parity test to gpurun_out/parity_observed.jsonl (brought back
    from the GPU box; the numbers quoted in DESIGN.md and profiles/ come from there)."""
    import json

    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_observed.jsonl"), "a") as fh:
            fh.write(json.dumps({"case": name, "observed": payload}, default=float) + "\n")
    except OSError:
        pass


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_greens():
    return load_golden("greens_kat.npz")


@pytest.fixture(scope="session")
def golden_spline():
    return load_golden("spline_small.npz")


@pytest.fixture(scope="session")
def golden_vector():
    return load_golden("vector_small.npz")


@pytest.fixture(scope="session")
def golden_cv():
    return load_golden("splinecv_small.npz")


@pytest.fixture(scope="session")
def golden_cfg1():
    return load_golden("cfg1_5k.npz")


@pytest.fixture(scope="session")
def oracle():
    import verde_oracle

    return verde_oracle


def greens_tolerance(d, scale=1.0):
    """
    Element-wise tolerance for scalar Green's values g = d^2 (log d - 1).

    The bar is 1e-12 RELATIVE (BASELINE.json north_star).  It is decidable where
    the reference's own evaluation is that accurate; it is not
      * for d < 1, where the reference evaluates d*(log(d**d) - d) with relative
        noise ~eps/(d |log d|): absolute bound 1e-15 there;
      * in the thin annulus |log d - 1| < 1e-3 around d = e, where g passes through
        zero and ANY 1-ulp difference in log is amplified: bound 1e-12 * d^2 there.
    """
    d = np.asarray(d, dtype=np.float64)
    with np.errstate(all="ignore"):
        g = np.abs(d * d * (np.log(np.maximum(d, 1e-300)) - 1.0))
        near_e = np.abs(np.log(np.maximum(d, 1e-300)) - 1.0) < 1e-3
    tol = 1e-12 * g
    tol = np.where(near_e, 1e-12 * d * d, tol)
    tol = np.where(d < 1, np.maximum(tol, 1e-15), tol)
    return tol * scale
