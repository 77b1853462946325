"""
Config-shaped golden fixtures from the UNMODIFIED reference (BASELINE configs 3, 4, 5 at sizes the
reference finishes in minutes on 8 cores) and the ``damping=None`` branch.

Run in the build container only (needs /root/reference):

    python oracle/make_golden_configs.py [cfg3] [cfg4] [cfg5] [none]

TEST INFRASTRUCTURE ONLY.  Every file stores the forces, a sub-sampled grid, the condition number of
the scaled damped normal matrix the reference solved (``numpy.linalg.eigvalsh``) and the library
versions; inputs are regenerated in the tests from the same seeds (heads of the arrays are stored to
prove the generators agree).

* ``cfg3_40k.npz``  -- config 3's shape: 40 000 CheckerBoard points, forces at the
  ``BlockReduce(np.mean, shape=(71, 71), center_coordinates=True)`` block centres (<= 5041), weights
  ``RandomState(1).uniform(0.1, 10)``; damping 1e-6 (as BASELINE) and a well-conditioned twin (1e-2).
* ``cfg4_3k.npz``   -- config 4's shape: ``VectorSpline2D(poisson=0.5, mindist=10e3, damping=1e-3)`` on 3000
  stations (6000 x 6000 coupled system), the synthetic two-component field of SURVEY 8d.
* ``cfg5_2k.npz``   -- config 5's shape: ``SplineCV`` 6 dampings x 4 mindists, default 5-fold CV, 2000 points.
* ``damping_none.npz`` -- ``Spline()`` / ``Spline(force_coords=...)`` / ``VectorSpline2D(damping=None)`` (the
  reference default): scikit-learn ``LinearRegression`` -> ``scipy.linalg.lstsq(cond=tol)``; the singular
  values of the scaled Jacobian and the rank LAPACK kept are stored so that the truncation is decidable.
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refimport import import_reference  # noqa: E402
from make_golden import GOLDEN, versions  # noqa: E402


def normal_cond(jac, weights, damping):
    "cond_2 of Js^T W Js + damping I with Js the StandardScaler-scaled Jacobian (what Ridge solves)."
    from sklearn.preprocessing import StandardScaler

    js = StandardScaler(with_mean=False).fit_transform(jac.copy())
    if weights is not None:
        js *= np.sqrt(weights)[:, None]
    a = js.T @ js
    a.flat[:: a.shape[0] + 1] += damping
    ev = np.linalg.eigvalsh(a)
    return float(ev[-1] / max(ev[0], 1e-300)), float(ev[-1])


def cfg3(vd, ver):
    t0 = time.time()
    cb = vd.synthetic.CheckerBoard()
    tab = cb.scatter(size=40_000, random_state=0)
    e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
    w = np.random.RandomState(1).uniform(0.1, 10, e.size)
    reducer = vd.BlockReduce(np.mean, shape=(71, 71), center_coordinates=True)
    (fe, fn), _ = reducer.filter((e, n), d)
    full = vd.grid_coordinates((e.min(), e.max(), n.min(), n.max()), shape=(600, 600))
    sub = (full[0][::8, ::8], full[1][::8, ::8])
    out = dict(east_head=e[:16], north_head=n[:16], data_head=d[:16], weights_head=w[:16],
               force_east=fe, force_north=fn, n_forces=fe.size)
    jac = vd.Spline(mindist=1e-5).jacobian((e, n), (fe, fn))
    for tag, damping in (("d1em6", 1e-6), ("d1em2", 1e-2)):
        sp = vd.Spline(damping=damping, mindist=1e-5, force_coords=(fe, fn)).fit((e, n), d, weights=w)
        out["force_" + tag] = sp.force_
        out["grid_sub_" + tag] = sp.predict(sub)
        out["pred_data_" + tag] = sp.predict((e[:500], n[:500]))
        out["cond_" + tag], out["lambda_max_" + tag] = normal_cond(jac, w, damping)
        print("cfg3", tag, "forces", fe.size, "cond %.3e" % out["cond_" + tag], "%.0f s" % (time.time() - t0))
    np.savez_compressed(os.path.join(GOLDEN, "cfg3_40k.npz"), **out, **ver)


def vector_field(se, sn):
    ue = 10 * np.sin(2 * np.pi * se / 3e5) * np.cos(2 * np.pi * sn / 4.5e5) + 3
    un = -7 * np.cos(2 * np.pi * se / 4.5e5) * np.sin(2 * np.pi * sn / 3e5) + 1
    return ue, un


def cfg4(vd, ver):
    t0 = time.time()
    region = (0, 500e3, 0, 500e3)
    se, sn = vd.scatter_points(region, 3000, random_state=0)
    ue, un = vector_field(se, sn)
    full = vd.grid_coordinates(region, shape=(200, 200))
    sub = (full[0][::5, ::5], full[1][::5, ::5])
    out = dict(east_head=se[:16], north_head=sn[:16], data_east_head=ue[:16], data_north_head=un[:16])
    for tag, damping in (("d1em3", 1e-3),):
        vs = vd.VectorSpline2D(poisson=0.5, mindist=10e3, damping=damping).fit((se, sn), (ue, un))
        out["force_" + tag] = vs.force_
        ge, gn = vs.predict(sub)
        out["grid_east_sub_" + tag], out["grid_north_sub_" + tag] = ge, gn
        out["score_" + tag] = vs.score((se, sn), (ue, un))
        jac = vs.jacobian((se, sn), (se, sn))
        out["cond_" + tag], out["lambda_max_" + tag] = normal_cond(jac, None, damping)
        print("cfg4", tag, "cond %.3e" % out["cond_" + tag], "%.0f s" % (time.time() - t0))
    np.savez_compressed(os.path.join(GOLDEN, "cfg4_3k.npz"), **out, **ver)


def cfg5(vd, ver):
    t0 = time.time()
    cb = vd.synthetic.CheckerBoard()
    tab = cb.scatter(size=2000, random_state=0)
    e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
    dampings = [1e-8, 1e-6, 1e-4, 1e-3, 1e-2, 1e-1]
    mindists = [0, 1e-5, 1e-3, 1e-1]
    cv = vd.SplineCV(dampings=dampings, mindists=mindists).fit((e, n), d)
    full = vd.grid_coordinates((e.min(), e.max(), n.min(), n.max()), shape=(120, 120))
    sub = (full[0][::3, ::3], full[1][::3, ::3])
    jac = cv.spline_.jacobian((e, n), (e, n))
    cond, lam = normal_cond(jac, None, cv.damping_)
    out = dict(east_head=e[:16], data_head=d[:16], dampings=np.array(dampings), mindists=np.array(mindists, dtype=float),
               scores=np.asarray(cv.scores_), best_damping=cv.damping_, best_mindist=cv.mindist_, force=cv.force_,
               grid_sub=cv.predict(sub), cond_best=cond, lambda_max_best=lam)
    print("cfg5 best", cv.mindist_, cv.damping_, "cond %.3e" % cond, "%.0f s" % (time.time() - t0))
    np.savez_compressed(os.path.join(GOLDEN, "cfg5_2k.npz"), **out, **ver)


def damping_none(vd, ver):
    """
    verde/base/least_squares.py:65-70 with damping=None: LinearRegression(fit_intercept=False).fit(Js, d, w)
    -> scipy.linalg.lstsq(sqrt(w) Js, sqrt(w) d, cond=LinearRegression.tol) (sklearn/linear_model/_base.py).
    """
    from sklearn.linear_model import LinearRegression
    from sklearn.preprocessing import StandardScaler

    cb = vd.synthetic.CheckerBoard()
    out = dict(lstsq_cond=float(LinearRegression().tol))
    for tag, size in (("n240", 240), ("n1500", 1500)):
        tab = cb.scatter(size=size, random_state=0)
        e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
        grid = vd.grid_coordinates(cb.region, shape=(23, 29))
        sp = vd.Spline().fit((e, n), d)
        js = StandardScaler(with_mean=False).fit_transform(sp.jacobian((e, n), (e, n)))
        sv = np.linalg.svd(js, compute_uv=False)
        regr = LinearRegression(fit_intercept=False).fit(js, d)
        out["east_head_" + tag], out["data_head_" + tag] = e[:16], d[:16]
        out["force_" + tag], out["grid_" + tag] = sp.force_, sp.predict(grid)
        out["singular_" + tag], out["rank_" + tag] = sv, int(regr.rank_)
        out["score_" + tag] = sp.score((e, n), d)
        print("none", tag, "rank", regr.rank_, "of", size, "sv max/min %.3e" % (sv[0] / sv[-1]))
    # separate forces (over-determined), weighted: the case the reference's warning is about
    tab = cb.scatter(size=1500, random_state=0)
    e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
    w = np.random.RandomState(1).uniform(0.1, 10, e.size)
    fc = vd.scatter_points(cb.region, 400, random_state=3)
    grid = vd.grid_coordinates(cb.region, shape=(23, 29))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sp = vd.Spline(force_coords=fc).fit((e, n), d, weights=w)
    js = StandardScaler(with_mean=False).fit_transform(sp.jacobian((e, n), fc)) * np.sqrt(w)[:, None]
    sv = np.linalg.svd(js, compute_uv=False)
    out["fc_east"], out["fc_north"] = fc
    out["force_fc_w"], out["grid_fc_w"] = sp.force_, sp.predict(grid)
    out["singular_fc_w"] = sv
    out["rank_fc_w"] = int(np.sum(sv > out["lstsq_cond"] * sv[0]))
    print("none fc_w rank", out["rank_fc_w"], "of", fc[0].size)
    # under-determined: 50 data, 240 forces (minimum-norm solution of the scaled system)
    tab = cb.scatter(size=240, random_state=0)
    e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sp = vd.Spline(force_coords=(e, n)).fit((e[:50], n[:50]), d[:50])
    out["force_under"], out["grid_under"] = sp.force_, sp.predict(grid)
    # elastic spline with the reference default damping=None
    region = (0, 500e3, 0, 500e3)
    se, sn = vd.scatter_points(region, 150, random_state=0)
    ue, un = vector_field(se, sn)
    vgrid = vd.grid_coordinates(region, shape=(11, 13))
    vs = vd.VectorSpline2D().fit((se, sn), (ue, un))
    out["force_vector"] = vs.force_
    out["grid_east_vector"], out["grid_north_vector"] = vs.predict(vgrid)
    js = StandardScaler(with_mean=False).fit_transform(vs.jacobian((se, sn), (se, sn)))
    sv = np.linalg.svd(js, compute_uv=False)
    out["singular_vector"] = sv
    out["rank_vector"] = int(np.sum(sv > out["lstsq_cond"] * sv[0]))
    print("none vector rank", out["rank_vector"], "of", 2 * se.size)
    np.savez_compressed(os.path.join(GOLDEN, "damping_none.npz"), **out, **ver)


def main():
    warnings.simplefilter("ignore")
    vd = import_reference()
    ver = versions()
    which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5", "none"]
    for name in which:
        {"cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5, "none": damping_none}[name](vd, ver)
    for name in sorted(os.listdir(GOLDEN)):
        print(name, os.path.getsize(os.path.join(GOLDEN, name)))


if __name__ == "__main__":
    main()
