"""
GPU parity of the Green's-function kernels (K0 Jacobian, K3 predict) through the
C ABI, against golden vectors produced by the unmodified reference and against
the CPU oracle on seeded inputs.
"""
import numpy as np
import pytest

from conftest import greens_tolerance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kernels():
    import verde_b200.kernels as k

    return k


def _dist(e, n, fe, fn, mindist):
    return np.sqrt((e[:, None] - fe[None, :]) ** 2 + (n[:, None] - fn[None, :]) ** 2) + mindist


@pytest.mark.parametrize("tag,mindist", [("md0", 0.0), ("md1em5", 1e-5)])
def test_jacobian_matches_reference_golden(kernels, golden_spline, tag, mindist):
    g = golden_spline
    e, n = g["east"], g["north"]
    jac = np.empty((e.size, e.size))
    kernels.jacobian_b200(e, n, e, n, mindist, jac)
    ref = g["jac_" + tag]
    tol = greens_tolerance(_dist(e, n, e, n, mindist))
    err = np.abs(jac - ref)
    assert np.all(err <= tol), "max err/tol = {}".format(np.max(err / np.maximum(tol, 1e-300)))
    # the diagonal (d = mindist) must be exactly 0 for mindist = 0, as in the reference
    if mindist == 0.0:
        assert np.all(np.diag(jac) == 0.0)


def test_jacobian_known_answers(kernels, golden_greens):
    """d = 0, d < 1, d = 1, d ~ e, huge d: the reference's greens_func_jit outputs."""
    g = golden_greens
    e, n = g["east"], g["north"]
    zero = np.zeros(1)
    for tag, mindist in (("md0", 0.0), ("md1em5", 1e-5), ("md10", 10.0)):
        jac = np.empty((e.size, 1))
        kernels.jacobian_b200(e, n, zero, zero, mindist, jac)
        ref = g["jit_" + tag]
        d = np.sqrt(e**2 + n**2) + mindist
        tol = greens_tolerance(d)
        err = np.abs(jac[:, 0] - ref)
        assert np.all(err <= tol), (tag, np.max(err / np.maximum(tol, 1e-300)))
    jac = np.empty((1, 1))
    kernels.jacobian_b200(zero, zero, zero, zero, 0.0, jac)
    assert jac[0, 0] == 0.0  # d == 0 -> exactly 0 (verde/spline.py:601-602)
    one = np.ones(1)
    kernels.jacobian_b200(one, zero, zero, zero, 0.0, jac)
    assert jac[0, 0] == -1.0  # d == 1 takes the big branch: 1*(log 1 - 1)


def test_jacobian_rectangular_vs_oracle(kernels, oracle):
    rng = np.random.RandomState(11)
    e, n = rng.uniform(0, 5000, 777), rng.uniform(-5000, 0, 777)
    fe, fn = rng.uniform(0, 5000, 301), rng.uniform(-5000, 0, 301)
    jac = np.empty((777, 301))
    kernels.jacobian_b200(e, n, fe, fn, 1e-5, jac)
    ref = oracle.jacobian(e, n, fe, fn, 1e-5)
    tol = greens_tolerance(_dist(e, n, fe, fn, 1e-5))
    assert np.all(np.abs(jac - ref) <= tol)


def test_jacobian_2d_matches_reference_golden(kernels, golden_vector):
    g = golden_vector
    e, n = g["east"], g["north"]
    jac = np.empty((2 * e.size, 2 * e.size))
    kernels.jacobian_2d_b200(e, n, e, n, 10e3, 0.5, jac)
    ref = g["jac_md1e4_p05"]
    # entries are O(10): log term ~ 2.5*log(1e4..7e5) plus O(1) ratios; 1e-12 relative to that scale
    np.testing.assert_allclose(jac, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())


def test_jacobian_2d_known_answers(kernels, golden_greens):
    g = golden_greens
    e, n = g["east_2d"], g["north_2d"]
    zero = np.zeros(1)
    for tag, mindist, poisson in (("md1e4_p05", 10e3, 0.5), ("md0_p025", 0.0, 0.25), ("md100_p0", 100.0, 0.0)):
        jac = np.empty((2 * e.size, 2))
        kernels.jacobian_2d_b200(e, n, zero, zero, mindist, poisson, jac)
        ne = e.size
        scale = np.maximum(np.abs(g["ee_" + tag]), np.abs(g["nn_" + tag])) + 1.0
        for got, name in ((jac[:ne, 0], "ee"), (jac[ne:, 1], "nn"), (jac[:ne, 1], "ne"), (jac[ne:, 0], "ne")):
            ref = g[name + "_" + tag]
            assert np.all(np.abs(got - ref) <= 1e-12 * scale), (tag, name)


def _predict_tolerance(oracle, e, n, fe, fn, mindist, f, rel=1e-12):
    """1e-12 relative to sum_j |g_j f_j| -- the sums cancel by 1e4..1e6."""
    jac = np.abs(oracle.jacobian(np.ravel(e), np.ravel(n), fe, fn, mindist))
    return rel * (jac @ np.abs(f))


@pytest.mark.parametrize("exact", [0, 1])
def test_predict_matches_reference_golden(kernels, golden_spline, oracle, exact):
    import verde_b200._cabi as cabi

    g = golden_spline
    e, n, f = g["east"], g["north"], g["force_pinned"]
    ge, gn = g["grid_east"], g["grid_north"]
    cabi.load().vb200_set_option(b"predict_exact", float(exact))
    try:
        out = np.empty(ge.size)
        kernels.predict_b200(ge.ravel(), gn.ravel(), e, n, 1e-5, f, out)
    finally:
        cabi.load().vb200_set_option(b"predict_exact", 0.0)
    ref = g["predict_pinned_md1em5"]
    tol = _predict_tolerance(oracle, ge, gn, e, n, 1e-5, f)
    err = np.abs(out - ref)
    assert np.all(err <= tol), np.max(err / tol)


@pytest.mark.parametrize("npts,m,mindist", [(1, 1, 0.0), (5, 3, 0.0), (1000, 70, 0.0), (4099, 2500, 1e-5),
                                            (300, 9000, 0.5), (70000, 33, 0.0)])
def test_predict_vs_oracle_shapes(kernels, oracle, npts, m, mindist):
    """ragged sizes: fewer forces than lanes, force splits, points not a multiple of the CTA tile"""
    rng = np.random.RandomState(npts + m)
    e, n = rng.uniform(0, 5000, npts), rng.uniform(-5000, 0, npts)
    fe, fn, f = rng.uniform(0, 5000, m), rng.uniform(-5000, 0, m), rng.normal(size=m)
    if m > 2:
        e[: min(npts, m) // 2] = fe[: min(npts, m) // 2]  # coincident points: d = mindist
        n[: min(npts, m) // 2] = fn[: min(npts, m) // 2]
    out = np.empty(npts)
    kernels.predict_b200(e, n, fe, fn, mindist, f, out)
    ref = oracle.predict(e, n, fe, fn, mindist, f)
    tol = _predict_tolerance(oracle, e, n, fe, fn, mindist, f) + 1e-300
    err = np.abs(out - ref)
    assert np.all(err <= tol), np.max(err / tol)


def test_predict_empty_inputs(kernels):
    out = np.empty(0)
    kernels.predict_b200(np.empty(0), np.empty(0), np.zeros(3), np.zeros(3), 0.0, np.ones(3), out)
    out = np.full(4, 7.0)
    kernels.predict_b200(np.arange(4.0), np.arange(4.0), np.empty(0), np.empty(0), 0.0, np.empty(0), out)
    assert np.all(out == 0.0)  # no forces: the reference's loop leaves result[i] = 0


def test_predict_linearity(kernels):
    """size-independent property at a larger size: predict is linear in the forces"""
    rng = np.random.RandomState(3)
    npts, m = 20000, 3000
    e, n = rng.uniform(0, 5000, npts), rng.uniform(-5000, 0, npts)
    fe, fn = rng.uniform(0, 5000, m), rng.uniform(-5000, 0, m)
    f1, f2 = rng.normal(size=m), rng.normal(size=m)
    o1, o2, o3 = np.empty(npts), np.empty(npts), np.empty(npts)
    kernels.predict_b200(e, n, fe, fn, 1e-5, f1, o1)
    kernels.predict_b200(e, n, fe, fn, 1e-5, f2, o2)
    kernels.predict_b200(e, n, fe, fn, 1e-5, 2.0 * f1 - 3.0 * f2, o3)
    scale = np.abs(o1).max() + np.abs(o2).max()
    np.testing.assert_allclose(o3, 2.0 * o1 - 3.0 * o2, rtol=0, atol=1e-9 * scale)


@pytest.mark.parametrize("exact", [0, 1])
def test_predict_2d_vs_oracle(kernels, oracle, golden_vector, exact):
    import verde_b200._cabi as cabi

    g = golden_vector
    e, n = g["east"], g["north"]
    f = g["force_d1em3"]
    ge, gn = g["grid_east"], g["grid_north"]
    cabi.load().vb200_set_option(b"predict_exact", float(exact))
    try:
        oe, on = np.empty(ge.size), np.empty(ge.size)
        kernels.predict_2d_b200(ge.ravel(), gn.ravel(), e, n, 10e3, 0.5, f, oe, on)
    finally:
        cabi.load().vb200_set_option(b"predict_exact", 0.0)
    jac = np.abs(oracle.jacobian_2d(ge.ravel(), gn.ravel(), e, n, 10e3, 0.5))
    tol = 1e-12 * (jac @ np.abs(f))
    ne = ge.size
    assert np.all(np.abs(oe - g["grid_east_d1em3"].ravel()) <= tol[:ne])
    assert np.all(np.abs(on - g["grid_north_d1em3"].ravel()) <= tol[ne:])


@pytest.mark.parametrize("ncomp,npts,m,mindist", [(2, 70_000, 3000, 0.0), (3, 70_000, 3000, 1e-5), (2, 33, 7, 0.0),
                                                  (4, 5000, 40_000, 1e-5), (5, 999, 1200, 0.0)])
def test_predict_multi_equals_one_predict_per_component(kernels, ncomp, npts, m, mindist):
    """vb200_predict_multi: several coefficient vectors on shared forces, Green's function evaluated once per
    pair; same lane striding / chunking / force split as vb200_predict -> every component bit-identical."""
    rng = np.random.RandomState(ncomp * 1000 + m)
    e, n = rng.uniform(0, 5000, npts), rng.uniform(-5000, 0, npts)
    fe, fn = rng.uniform(0, 5000, m), rng.uniform(-5000, 0, m)
    forces = rng.normal(size=(ncomp, m))
    multi = kernels.predict_multi_b200(e, n, fe, fn, mindist, forces, np.empty((ncomp, npts)))
    for k in range(ncomp):
        single = kernels.predict_b200(e, n, fe, fn, mindist, forces[k], np.empty(npts))
        np.testing.assert_array_equal(multi[k], single)


def test_predict_multi_non_finite_inputs_stay_in_their_component(kernels):
    rng = np.random.RandomState(3)
    e, n = rng.uniform(0, 10, 300), rng.uniform(0, 10, 300)
    fe, fn = rng.uniform(0, 10, 50), rng.uniform(0, 10, 50)
    forces = rng.normal(size=(2, 50))
    forces[1, 7] = np.nan
    e[5] = np.inf
    out = kernels.predict_multi_b200(e, n, fe, fn, 0.0, forces, np.empty((2, 300)))
    assert np.isnan(out[1]).all()                         # a NaN coefficient poisons its own component ...
    assert np.isnan(out[0, 5]) and np.isfinite(np.delete(out[0], 5)).all()  # ... and only that one
