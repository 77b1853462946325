"""
GPU parity at the SHAPES of BASELINE configs 2-5.

* configs 3, 4, 5 at sizes the unmodified reference finishes in minutes (goldens from
  oracle/make_golden_configs.py: 40 000 points x 5 041 BlockReduce-centre forces, weighted; VectorSpline2D on
  3 000 stations = 6 000 x 6 000 coupled system; SplineCV 6 dampings x 4 mindists x 5 folds on 2 000 points);
* configs 2 and 4 at FULL size (50 000 forces; 20 000 stations = 40 000 unknowns) through what the library
  itself reports per fit -- the matrix-free backward error of the normal equations and the condition estimate --
  cross-checked once against an independent evaluation.

Tolerances carry cond_2 of the matrix the reference factored (stored in the golden): forces to
max(1e-9, 10 cond eps), gridded values to max(1e-6, 1e-3 cond eps) of the data range; every assertion message
states cond.  Observed gaps are appended to gpurun_out/parity_observed.jsonl.
"""
import warnings

import numpy as np
import pytest
from conftest import load_golden, record_parity

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module")
def vb():
    import verde_b200

    return verde_b200


def force_tol(cond):
    return max(1e-9, 10 * cond * EPS)


def grid_tol(cond):
    return max(1e-6, 1e-3 * cond * EPS)


def relmax(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def test_config3_shape_40k_points_blockreduce_forces_weighted(vb):
    """Config 3's shape: data rows >> forces, forces at BlockReduce centres, weights (verde/spline.py:426-464)."""
    g = load_golden("cfg3_40k.npz")
    tab = vb.CheckerBoard().scatter(size=40_000, random_state=0)
    e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
    w = np.random.RandomState(1).uniform(0.1, 10, e.size)
    np.testing.assert_array_equal(e[:16], g["east_head"])
    np.testing.assert_array_equal(d[:16], g["data_head"])
    np.testing.assert_array_equal(w[:16], g["weights_head"])
    # the force positions come from the device BlockReduce, as in the config; they must be the reference's
    (fe, fn), _ = vb.BlockReduce(np.mean, shape=(71, 71), center_coordinates=True).filter((e, n), d)
    assert fe.size == int(g["n_forces"]) == g["force_east"].size
    np.testing.assert_allclose(fe, g["force_east"], rtol=1e-13, atol=1e-9)
    np.testing.assert_allclose(fn, g["force_north"], rtol=1e-13, atol=1e-9)
    fc = (g["force_east"], g["force_north"])
    full = vb.grid_coordinates((e.min(), e.max(), n.min(), n.max()), shape=(600, 600))
    sub = (full[0][::8, ::8], full[1][::8, ::8])
    rng = np.ptp(d)
    observed = {}
    for tag, damping in (("d1em6", 1e-6), ("d1em2", 1e-2)):
        cond = float(g["cond_" + tag])  # 2.7e15 and 1.9e11
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", FutureWarning)
            sp = vb.Spline(damping=damping, mindist=1e-5, force_coords=fc).fit((e, n), d, weights=w)
        ferr = relmax(sp.force_, g["force_" + tag])
        gerr = float(np.max(np.abs(sp.predict(sub) - g["grid_sub_" + tag])) / rng)
        derr = float(np.max(np.abs(sp.predict((e[:500], n[:500])) - g["pred_data_" + tag])) / rng)
        info = sp.fit_info_
        observed[tag] = dict(cond=cond, force=ferr, grid=gerr, data=derr, cond_est=info["cond_est"],
                             cond_upper=info["cond_upper"], backward_err=info["backward_err"])
        msg = "cfg3-shape {}: cond={:.2e} {}".format(tag, cond, observed[tag])
        assert ferr <= force_tol(cond), msg
        assert gerr <= grid_tol(cond) and derr <= grid_tol(cond), msg
        assert info["cond_est"] <= 2 * cond and info["cond_upper"] >= 0.5 * cond, msg
        assert 0 <= info["backward_err"] < 1e-13, msg
        assert abs(info["lambda_max"] / float(g["lambda_max_" + tag]) - 1) < 1e-2, msg
    record_parity("cfg3_40k", observed)


def vector_field(se, sn):
    ue = 10 * np.sin(2 * np.pi * se / 3e5) * np.cos(2 * np.pi * sn / 4.5e5) + 3
    un = -7 * np.cos(2 * np.pi * se / 4.5e5) * np.sin(2 * np.pi * sn / 3e5) + 1
    return ue, un


def test_config4_shape_vector_spline_3000_stations(vb):
    """Config 4's shape: VectorSpline2D(poisson=0.5, mindist=10e3, damping=1e-3), 6000 x 6000 coupled system."""
    g = load_golden("cfg4_3k.npz")
    region = (0, 500e3, 0, 500e3)
    se, sn = vb.scatter_points(region, 3000, random_state=0)
    ue, un = vector_field(se, sn)
    np.testing.assert_array_equal(se[:16], g["east_head"])
    np.testing.assert_array_equal(un[:16], g["data_north_head"])
    cond = float(g["cond_d1em3"])  # 3.6e10
    vs = vb.VectorSpline2D(poisson=0.5, mindist=10e3, damping=1e-3).fit((se, sn), (ue, un))
    full = vb.grid_coordinates(region, shape=(200, 200))
    sub = (full[0][::5, ::5], full[1][::5, ::5])
    ge, gn = vs.predict(sub)
    ferr = relmax(vs.force_, g["force_d1em3"])
    gerr = max(float(np.max(np.abs(ge - g["grid_east_sub_d1em3"])) / np.ptp(ue)),
               float(np.max(np.abs(gn - g["grid_north_sub_d1em3"])) / np.ptp(un)))
    score = vs.score((se, sn), (ue, un))
    info = vs.fit_info_
    observed = dict(cond=cond, force=ferr, grid=gerr, score_diff=abs(score - float(g["score_d1em3"])),
                    cond_est=info["cond_est"], cond_upper=info["cond_upper"], backward_err=info["backward_err"])
    record_parity("cfg4_3k", observed)
    msg = "cfg4-shape: cond={:.2e} {}".format(cond, observed)
    assert ferr <= force_tol(cond), msg
    assert gerr <= 1e-6, msg
    assert observed["score_diff"] <= 1e-9, msg
    assert info["cond_est"] <= 2 * cond and info["cond_upper"] >= 0.5 * cond, msg
    assert 0 <= info["backward_err"] < 1e-13, msg


def test_config5_shape_splinecv_6x4x5_on_2000_points(vb):
    """Config 5's shape: 6 dampings x 4 mindists, default 5-fold CV (verde/spline.py:183-264); scores_ is
    mindist-major.  The reference's 120 independent fits vs 20 Gram matrices factored 6 at a time here."""
    g = load_golden("cfg5_2k.npz")
    tab = vb.CheckerBoard().scatter(size=2000, random_state=0)
    e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
    np.testing.assert_array_equal(e[:16], g["east_head"])
    dampings, mindists = list(g["dampings"]), list(g["mindists"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", FutureWarning)
        cv = vb.SplineCV(dampings=dampings, mindists=mindists).fit((e, n), d)
    ref = g["scores"]
    assert cv.scores_.shape == ref.shape == (24,)
    diff = np.abs(cv.scores_ - ref)
    # R^2 of a fold's predictions inherits the solve's cond * eps: systems with damping 1e-8 sit at cond ~ 1e15
    # (2000 points), those with damping >= 1e-4 at cond <= 1e11
    # (observed on B200: 2.1e-8 at damping 1e-8, 1.3e-9 at 1e-6, <= 2.1e-10 from 1e-4 up)
    by_damping = {1e-8: 1e-6, 1e-6: 1e-7}
    tol = np.array([by_damping.get(float(damp), 1e-8) for _ in mindists for damp in dampings])
    observed = dict(score_max_diff=float(diff.max()), score_diff_by_damping={
        str(damp): float(diff.reshape(len(mindists), len(dampings))[:, k].max()) for k, damp in enumerate(dampings)},
        best=(float(cv.mindist_), float(cv.damping_)), cond_best=float(g["cond_best"]))
    assert np.all(diff <= tol), observed
    assert (cv.mindist_, cv.damping_) == (float(g["best_mindist"]), float(g["best_damping"])), observed
    assert int(np.argmax(cv.scores_)) == int(np.argmax(ref))
    cond = float(g["cond_best"])  # 7.4e14: the winner is the smallest damping
    full = vb.grid_coordinates((e.min(), e.max(), n.min(), n.max()), shape=(120, 120))
    sub = (full[0][::3, ::3], full[1][::3, ::3])
    gerr = float(np.max(np.abs(cv.predict(sub) - g["grid_sub"])) / np.ptp(d))
    ferr = relmax(cv.force_, g["force"])
    observed.update(grid=gerr, force=ferr)
    record_parity("cfg5_2k", observed)
    assert gerr <= grid_tol(cond) and ferr <= force_tol(cond), "cfg5-shape final fit: cond={:.2e} {}".format(cond, observed)


def test_config2_full_size_50k_backward_error(vb):
    """
    BASELINE config 2 itself: 50 000 CheckerBoard points, forces at the data, damping 1e-6, mindist 1e-5.
    No CPU reference exists at this size (the reference needs ~80 GB and hours); what can be decided is whether
    the forces solve the reference's normal equations: the library's matrix-free backward error, confirmed by an
    independent evaluation on a random subset of equations with the K0 Jacobian kernel + torch (checker only).
    """
    import torch

    from verde_b200 import _cabi

    lib = _cabi.require_device()
    n, alpha, mindist = 50_000, 1e-6, 1e-5
    tab = vb.CheckerBoard().scatter(size=n, random_state=0)
    e, nn, d = tab.easting.values, tab.northing.values, tab.scalars.values
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", FutureWarning)
        sp = vb.Spline(damping=alpha, mindist=mindist).fit((e, nn), d)
    info = sp.fit_info_
    observed = {k: info[k] for k in ("cond_est", "cond_upper", "lambda_max", "lambda_min_est", "backward_err", "diag_ms",
                                     "gram_ms", "factor_ms", "solve_ms")}
    msg = "config 2 (50k): {}".format(observed)
    assert info["info"] == 0 and info["diag_min"] > 0, msg
    assert 0 <= info["backward_err"] < 1e-13, msg
    assert info["cond_upper"] >= info["cond_est"] > 1e12, msg  # cond ~ 1.85 n^2 / alpha ~ 4.6e15
    # independent check on 64 of the 50 000 normal equations: row j of Js^T (d - Js c) - alpha c
    dev = torch.device("cuda", torch.cuda.current_device())
    de, dn = torch.from_numpy(e).to(dev), torch.from_numpy(nn).to(dev)
    f, dd = torch.from_numpy(sp.force_).to(dev), torch.from_numpy(d).to(dev)
    pred = torch.from_numpy(sp.predict((e, nn))).to(dev)  # fast predict kernel: J f (1e-12 of sum |g f|)
    pick = np.random.RandomState(3).choice(n, 64, replace=False)
    cols_e, cols_n = de[pick].contiguous(), dn[pick].contiguous()
    jac = torch.empty((n, 64), dtype=torch.float64, device=dev)  # the picked COLUMNS of J
    assert lib.vb200_jacobian(_cabi.ptr(de), _cabi.ptr(dn), n, _cabi.ptr(cols_e), _cabi.ptr(cols_n), 64, mindist,
                              _cabi.ptr(jac)) == 0
    s = jac.std(dim=0, unbiased=False)
    c = f[pick] * s
    resid = (jac.T @ (dd - pred)) / s - alpha * c
    scale = (jac.abs().T @ (dd.abs() + pred.abs())) / s + alpha * c.abs()
    rowwise = float((resid.abs() / scale).max())
    observed["rowwise_backward_err_64_equations"] = rowwise
    record_parity("config2_50k_full", observed)
    # the residual d - J f cancels ~6 digits and J f itself is summed from terms 1e9 times larger: the row-wise
    # ratio sits at the rounding level of that evaluation, not at eps
    assert rowwise < 1e-9, msg + " rowwise {}".format(rowwise)
    assert np.sqrt(np.mean((pred.cpu().numpy()[:2000] - d[:2000]) ** 2)) < 2e-3 * np.ptp(d)


def test_config4_full_size_20k_stations_backward_error(vb):
    """BASELINE config 4 itself: VectorSpline2D on 20 000 stations (40 000 x 40 000 coupled system), damping 1e-3."""
    region = (0, 500e3, 0, 500e3)
    se, sn = vb.scatter_points(region, 20_000, random_state=0)
    ue, un = vector_field(se, sn)
    vs = vb.VectorSpline2D(poisson=0.5, mindist=10e3, damping=1e-3).fit((se, sn), (ue, un))
    info = vs.fit_info_
    observed = {k: info[k] for k in ("cond_est", "cond_upper", "lambda_max", "backward_err", "diag_ms", "gram_ms",
                                     "factor_ms", "solve_ms")}
    pe, pn = vs.predict((se[:3000], sn[:3000]))
    observed["misfit_east"] = float(np.max(np.abs(pe - ue[:3000])) / np.ptp(ue))
    observed["misfit_north"] = float(np.max(np.abs(pn - un[:3000])) / np.ptp(un))
    record_parity("config4_20k_stations_full", observed)
    msg = "config 4 (20k stations): {}".format(observed)
    assert info["info"] == 0 and info["cols"] == 40_000, msg
    assert 0 <= info["backward_err"] < 1e-13, msg
    assert info["cond_upper"] >= info["cond_est"] > 1e8, msg
    assert observed["misfit_east"] < 1e-2 and observed["misfit_north"] < 1e-2, msg


@pytest.mark.parametrize("tag,damping,mindist,weighted", [("d1em3", 1e-3, None, False), ("d1em6_md", 1e-6, 1e-5, False),
                                                          ("d1em8", 1e-8, None, False), ("d1em4_w", 1e-4, None, True)])
def test_condition_estimate_within_2x_of_eigvalsh(vb, tag, damping, mindist, weighted):
    """SURVEY 8b out_diag{cond}: the device estimate against numpy.linalg.eigvalsh of the matrix the reference
    factored (n = 240 goldens, cond 1e8 .. 1e13)."""
    g = load_golden("spline_small.npz")
    w = g["weights"] if weighted else None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", FutureWarning)
        sp = vb.Spline(damping=damping, mindist=mindist).fit((g["east"], g["north"]), g["data"], weights=w)
    cond = float(g["cond_" + tag])
    info = sp.fit_info_
    msg = "cond={:.3e} est={:.3e} upper={:.3e} lambda_max={:.3e} lambda_min_est={:.3e}".format(
        cond, info["cond_est"], info["cond_upper"], info["lambda_max"], info["lambda_min_est"])
    assert 0.5 * cond <= info["cond_est"] <= 2.0 * cond, msg
    assert info["cond_upper"] >= cond * (1 - 1e-6), msg
    assert 0 <= info["backward_err"] < 1e-14, msg


def test_diagnostics_can_be_switched_off(vb):
    from verde_b200 import _cabi

    lib = _cabi.require_device()
    g = load_golden("spline_small.npz")
    lib.vb200_set_option(b"fit_diagnostics", 0.0)
    try:
        sp = vb.Spline(damping=1e-3).fit((g["east"], g["north"]), g["data"])
        assert sp.fit_info_["cond_est"] == 0 and sp.fit_info_["diag_ms"] == 0
    finally:
        lib.vb200_set_option(b"fit_diagnostics", 1.0)
    ref = vb.Spline(damping=1e-3).fit((g["east"], g["north"]), g["data"])
    np.testing.assert_array_equal(sp.force_, ref.force_)  # diagnostics never touch the solution
    assert ref.fit_info_["cond_est"] > 0
