"""
GPU tests at sizes the CPU oracle cannot reach in seconds, through size-independent
properties (linearity, row-shard additivity, backward error of the solve, agreement of
SplineCV's Gram-reuse sweep with plain cross_val_score) -- all through the C ABI.
"""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch

    import verde_b200 as vb
    from verde_b200 import _cabi, kernels

    lib = _cabi.require_device()
    return vb, _cabi, kernels, lib, torch


def checkerboard_sample(n, seed=0):
    rng = np.random.RandomState(seed)
    e = rng.uniform(0, 5000, n)
    nn = rng.uniform(-5000, 0, n)
    d = 1000.0 * np.sin((2 * np.pi / 2500.0) * e) * np.cos((2 * np.pi / 2500.0) * nn)
    return e, nn, d


def test_row_sharded_accumulation_equals_single_pass(env):
    """
    The multi-GPU fit all-reduces per-rank partial normal equations.  Emulate 3 ranks on one
    GPU: accumulating three row ranges into the same buffers must give the same system (and
    the same forces) as one pass over all rows.
    """
    vb, _cabi, kernels, lib, torch = env
    n, m = 3001, 700
    e, nn, d = checkerboard_sample(n)
    w = np.random.RandomState(2).uniform(0.2, 4, n)
    fe, fn, _ = checkerboard_sample(m, seed=5)
    whole = kernels.GramSystem(m).accumulate(0, e, nn, d, w, fe, fn, 1e-5)
    parts = kernels.GramSystem(m)
    for r in range(3):
        a, b = vb.dist.shard_bounds(n, r, 3)
        parts.accumulate(0, e[a:b], nn[a:b], d[a:b], w[a:b], fe, fn, 1e-5, accumulate=(r > 0))
    g0, g1 = whole.gram.cpu().numpy(), parts.gram.cpu().numpy()
    s0, s1 = whole.stats.cpu().numpy(), parts.stats.cpu().numpy()
    np.testing.assert_allclose(g1, g0, rtol=1e-12, atol=1e-12 * np.abs(g0).max())
    np.testing.assert_allclose(s1, s0, rtol=1e-12, atol=1e-12 * np.abs(s0).max())
    f0 = whole.solve(n, 1e-4)
    f1 = parts.solve(n, 1e-4)
    ref = vb.Spline(damping=1e-4, mindist=None, force_coords=(fe, fn))
    ref.mindist = 1e-5
    ref.fit((e, nn), d, weights=w)
    np.testing.assert_array_equal(f0, ref.force_)  # staged API == fused API, bit for bit
    # forces agree to ~cond*eps (cond ~ 1e10 here); the predictions they imply agree far tighter
    assert np.max(np.abs(f1 - f0)) / np.max(np.abs(f0)) < 1e-4
    p0, p1 = np.empty(500), np.empty(500)
    kernels.predict_b200(e[:500], nn[:500], fe, fn, 1e-5, f0, p0)
    kernels.predict_b200(e[:500], nn[:500], fe, fn, 1e-5, f1, p1)
    assert np.max(np.abs(p1 - p0)) < 1e-6 * np.ptp(d)


def test_backward_error_of_the_solve_at_20k(env):
    """
    n = m = 20 000, damping 1e-6, mindist 1e-5 (config-2-like, cond ~ 7e14): the forces must
    satisfy the reference's normal equations  (D G D + alpha I) c = D b  to a normwise
    backward error of a few eps, checked with an independent FP64 evaluation
    (Jacobian from the K0 kernel, products by torch/cuBLAS -- checker only).
    """
    vb, _cabi, kernels, lib, torch = env
    n = 20000
    alpha, mindist = 1e-6, 1e-5
    e, nn, d = checkerboard_sample(n)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", FutureWarning)
        sp = vb.Spline(damping=alpha, mindist=mindist).fit((e, nn), d)
    dev = torch.device("cuda", torch.cuda.current_device())
    jac = torch.empty((n, n), dtype=torch.float64, device=dev)
    de, dn = torch.from_numpy(e).to(dev), torch.from_numpy(nn).to(dev)
    rc = lib.vb200_jacobian(_cabi.ptr(de), _cabi.ptr(dn), n, _cabi.ptr(de), _cabi.ptr(dn), n, mindist, _cabi.ptr(jac))
    assert rc == 0
    s = jac.std(dim=0, unbiased=False)
    f = torch.from_numpy(sp.force_).to(dev)
    dd = torch.from_numpy(d).to(dev)
    c = f * s                                    # scaled unknowns
    js_c = jac @ f                               # = Js c
    resid = (jac.T @ (dd - js_c)) / s - alpha * c  # D b - (D G D + alpha I) c
    bound = (jac.abs().T @ (jac.abs() @ f.abs())) / s + alpha * c.abs() + (jac.abs().T @ dd.abs()) / s
    backward = float((resid.abs() / bound).max())
    assert backward < 1e-13, backward
    # and the fit really interpolates: predictions at the data reproduce the data closely
    pred = sp.predict((e[:2000], nn[:2000]))
    assert np.sqrt(np.mean((pred - d[:2000]) ** 2)) < 2e-3 * np.ptp(d)
    assert sp.fit_info_["info"] == 0 and sp.fit_info_["diag_min"] > 0


def test_fit_is_linear_in_the_data(env):
    vb, _cabi, kernels, lib, torch = env
    n = 6000
    e, nn, d1 = checkerboard_sample(n)
    d2 = np.random.RandomState(9).normal(size=n)
    system = kernels.GramSystem(n)
    forces = []
    for data in (d1, d2, 2.0 * d1 - 0.5 * d2):
        system.accumulate(0, e, nn, data, None, e, nn, 0.0)
        forces.append(system.solve(n, 1e-2))
    comb = 2.0 * forces[0] - 0.5 * forces[1]
    assert np.max(np.abs(forces[2] - comb)) / np.max(np.abs(comb)) < 1e-7


def test_predict_grid_4m_points_consistency(env):
    """4M-point grid (the BASELINE grid size) against 2 000 forces: fast vs exact Green's
    function agree to 1e-12 of sum|g f| on a strided subset, and the result is independent of
    how the point range is split (the multi-GPU grid sharding)."""
    vb, _cabi, kernels, lib, torch = env
    m = 2000
    fe, fn, _ = checkerboard_sample(m, seed=3)
    f = np.random.RandomState(4).normal(size=m)
    gx, gy = vb.grid_coordinates((0, 5000, -5000, 0), shape=(2000, 2000))
    gx, gy = gx.ravel(), gy.ravel()
    out = np.empty(gx.size)
    kernels.predict_b200(gx, gy, fe, fn, 1e-5, f, out)
    sub = slice(0, gx.size, 997)
    lib.vb200_set_option(b"predict_exact", 1.0)
    try:
        exact = np.empty(gx[sub].size)
        kernels.predict_b200(gx[sub], gy[sub], fe, fn, 1e-5, f, exact)
    finally:
        lib.vb200_set_option(b"predict_exact", 0.0)
    d = np.sqrt((gx[sub][:, None] - fe[None, :]) ** 2 + (gy[sub][:, None] - fn[None, :]) ** 2) + 1e-5
    scale = (np.abs(d * d * (np.log(d) - 1)) @ np.abs(f))
    assert np.all(np.abs(out[sub] - exact) <= 1e-12 * scale)
    a, b = vb.dist.shard_bounds(gx.size, 3, 8)
    part = np.empty(b - a)
    kernels.predict_b200(gx[a:b], gy[a:b], fe, fn, 1e-5, f, part)
    np.testing.assert_allclose(part, out[a:b], rtol=0, atol=1e-12 * scale.max())


def test_splinecv_gram_reuse_equals_plain_cross_validation(env):
    vb, _cabi, kernels, lib, torch = env
    e, nn, d = checkerboard_sample(1500)
    dampings, mindists = [1e-5, 1e-3, 1e-1], [0, 1e-2]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", FutureWarning)
        cv = vb.SplineCV(dampings=dampings, mindists=mindists).fit((e, nn), d)
        plain = []
        for md in mindists:
            for damp in dampings:
                sp = vb.Spline(damping=damp)
                sp.mindist = md
                plain.append(np.mean(vb.cross_val_score(sp, (e, nn), d)))
    np.testing.assert_allclose(cv.scores_, plain, rtol=0, atol=1e-9)
    best = int(np.argmax(plain))
    assert (cv.mindist_, cv.damping_) == (mindists[best // 3], dampings[best % 3])


def test_vector_spline_sharded_accumulation(env):
    vb, _cabi, kernels, lib, torch = env
    rng = np.random.RandomState(1)
    n = 1200
    e, nn = rng.uniform(0, 500e3, n), rng.uniform(0, 500e3, n)
    ue = 10 * np.sin(2 * np.pi * e / 3e5) * np.cos(2 * np.pi * nn / 4.5e5) + 3
    un = -7 * np.cos(2 * np.pi * e / 4.5e5) * np.sin(2 * np.pi * nn / 3e5) + 1
    vs = vb.VectorSpline2D(damping=1e-3).fit((e, nn), (ue, un))
    system = kernels.GramSystem(2 * n)
    for r in range(2):
        a, b = vb.dist.shard_bounds(n, r, 2)
        system.accumulate(1, e[a:b], nn[a:b], np.concatenate([ue[a:b], un[a:b]]), None, e, nn, 10e3, 0.5,
                          accumulate=(r > 0))
    f = system.solve(2 * n, 1e-3)
    assert np.max(np.abs(f - vs.force_)) / np.max(np.abs(vs.force_)) < 1e-7
    pe, pn = vs.predict((e, nn))
    assert np.max(np.abs(pe - ue)) < 1e-2 * np.ptp(ue) and np.max(np.abs(pn - un)) < 1e-2 * np.ptp(un)
