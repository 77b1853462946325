"""
Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  The reference has no golden vectors of its own for
this path, so these outputs of the reference itself
are the pin for both the CPU oracle (tests/test_oracle.py) and the CUDA path
(tests/test_gpu_*.py).  Library versions are stored in every file because the
solve lives in un-vendored scikit-learn / scipy.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refimport import import_reference  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def versions():
    import numba
    import scipy
    import sklearn

    return dict(
        numpy_version=np.__version__,
        scipy_version=scipy.__version__,
        sklearn_version=sklearn.__version__,
        numba_version=numba.__version__,
    )


def cond_estimate(vd, jac, weights, damping):
    """cond_2 of the scaled damped normal matrix the reference solves."""
    from sklearn.preprocessing import StandardScaler

    js = StandardScaler(with_mean=False).fit_transform(jac.copy())
    if weights is not None:
        js = js * np.sqrt(weights)[:, None]
    a = js.T @ js
    a.flat[:: a.shape[0] + 1] += damping
    ev = np.linalg.eigvalsh(a)
    return float(ev[-1] / max(ev[0], 1e-300))


def main():
    warnings.simplefilter("ignore")
    vd = import_reference()
    from verde.spline import greens_func_jit, greens_func_numpy, jacobian_numba, predict_numba
    from verde.vector import GREENS_FUNC_2D_JIT, greens_func_2d

    os.makedirs(GOLDEN, exist_ok=True)
    ver = versions()

    # ------------------------------------------------------------------
    # 1. Green's function known-answer vectors (scalar + elastic)
    # ------------------------------------------------------------------
    rng = np.random.RandomState(42)
    special_d = np.array([0.0, 1e-300, 1e-12, 1e-5, 0.1, 0.5, 0.999999999, 1.0, 1.0000000001,
                          np.e * (1 - 1e-6), np.e, np.e * (1 + 1e-6), 10.0, 1234.5, 7071.07,
                          1e6, 1e9])
    east = np.concatenate([special_d, rng.uniform(-5000, 5000, 400), rng.uniform(-2, 2, 200)])
    north = np.concatenate([np.zeros_like(special_d), rng.uniform(-5000, 5000, 400),
                            rng.uniform(-2, 2, 200)])
    out = dict(east=east, north=north)
    for tag, mindist in (("md0", 0.0), ("md1em5", 1e-5), ("md10", 10.0)):
        out["jit_" + tag] = np.array([greens_func_jit(e, n, mindist) for e, n in zip(east, north)])
        out["numpy_" + tag] = greens_func_numpy(east.copy(), north.copy(), mindist)
    keep = (east**2 + north**2) > 0  # log(0) for mindist=0 in the elastic kernel
    e2, n2 = east[keep], north[keep]
    out["east_2d"], out["north_2d"] = e2, n2
    for tag, mindist, poisson in (("md1e4_p05", 10e3, 0.5), ("md0_p025", 0.0, 0.25),
                                  ("md100_p0", 100.0, 0.0)):
        ee, nn, ne = greens_func_2d(e2.copy(), n2.copy(), mindist, poisson)
        out["ee_" + tag], out["nn_" + tag], out["ne_" + tag] = ee, nn, ne
        jit = np.array([GREENS_FUNC_2D_JIT(a, b, mindist, poisson) for a, b in zip(e2, n2)])
        out["jit_ee_" + tag], out["jit_nn_" + tag], out["jit_ne_" + tag] = jit.T
    np.savez_compressed(os.path.join(GOLDEN, "greens_kat.npz"), **out, **ver)

    # ------------------------------------------------------------------
    # 2. Small scalar spline: Jacobian, forces (several dampings, weights,
    #    separate force_coords), predictions
    # ------------------------------------------------------------------
    cb = vd.synthetic.CheckerBoard()
    tab = cb.scatter(size=240, random_state=0)
    e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
    w = np.random.RandomState(1).uniform(0.1, 10, e.size)
    grid = vd.grid_coordinates(cb.region, shape=(17, 19))
    out = dict(east=e, north=n, data=d, weights=w, grid_east=grid[0], grid_north=grid[1])
    for tag, mindist in (("md0", 0.0), ("md1em5", 1e-5)):
        jac = np.empty((e.size, e.size))
        jacobian_numba(e, n, e, n, mindist, jac)
        out["jac_" + tag] = jac
    for tag, damping, mindist, weights in (
        ("d1em3", 1e-3, None, None),
        ("d1em6_md", 1e-6, 1e-5, None),
        ("d1em8", 1e-8, None, None),
        ("d1em4_w", 1e-4, None, w),
    ):
        sp = vd.Spline(damping=damping, mindist=mindist).fit((e, n), d, weights=weights)
        out["force_" + tag] = sp.force_
        out["grid_" + tag] = sp.predict(grid)
        out["cond_" + tag] = cond_estimate(vd, sp.jacobian((e, n), (e, n)), weights, damping)
        out["score_" + tag] = sp.score((e, n), d, weights=weights)
    # forces on a coarser set than the data (over-determined) and under-determined
    fc = vd.scatter_points(cb.region, 60, random_state=3)
    sp = vd.Spline(damping=1e-5, force_coords=fc).fit((e, n), d, weights=w)
    out["fc_east"], out["fc_north"] = fc
    out["force_fc"] = sp.force_
    out["grid_fc"] = sp.predict(grid)
    out["jac_fc"] = sp.jacobian((e, n), fc)
    sp = vd.Spline(damping=1e-4, force_coords=(e, n)).fit((e[:50], n[:50]), d[:50])
    out["force_under"] = sp.force_  # 50 data, 240 forces: sklearn takes the dual form
    out["grid_under"] = sp.predict(grid)
    # predict_numba with a pinned force vector (pure kernel arithmetic)
    fpin = np.random.RandomState(7).normal(size=e.size)
    res = np.empty(grid[0].size)
    predict_numba(grid[0].ravel(), grid[1].ravel(), e, n, 1e-5, fpin, res)
    out["force_pinned"] = fpin
    out["predict_pinned_md1em5"] = res.copy()
    np.savez_compressed(os.path.join(GOLDEN, "spline_small.npz"), **out, **ver)

    # ------------------------------------------------------------------
    # 3. Small VectorSpline2D
    # ------------------------------------------------------------------
    region = (0, 500e3, 0, 500e3)
    se, sn = vd.scatter_points(region, 150, random_state=0)
    ue = 10 * np.sin(2 * np.pi * se / 3e5) * np.cos(2 * np.pi * sn / 4.5e5) + 3
    un = -7 * np.cos(2 * np.pi * se / 4.5e5) * np.sin(2 * np.pi * sn / 3e5) + 1
    wv = (np.random.RandomState(2).uniform(0.5, 2, se.size),
          np.random.RandomState(3).uniform(0.5, 2, se.size))
    vgrid = vd.grid_coordinates(region, shape=(11, 13))
    out = dict(east=se, north=sn, data_east=ue, data_north=un, weights_east=wv[0],
               weights_north=wv[1], grid_east=vgrid[0], grid_north=vgrid[1])
    for tag, damping, mindist, poisson, weights in (
        ("d1em3", 1e-3, 10e3, 0.5, None),
        ("d1em6_w", 1e-6, 10e3, 0.5, wv),
        ("d1em2_p025", 1e-2, 5e3, 0.25, None),
    ):
        vs = vd.VectorSpline2D(poisson=poisson, mindist=mindist, damping=damping)
        vs.fit((se, sn), (ue, un), weights=weights)
        out["force_" + tag] = vs.force_
        ge, gn = vs.predict(vgrid)
        out["grid_east_" + tag], out["grid_north_" + tag] = ge, gn
        out["score_" + tag] = vs.score((se, sn), (ue, un), weights=weights)
        jac = vs.jacobian((se, sn), (se, sn))
        wcat = None if weights is None else np.concatenate(weights)
        out["cond_" + tag] = cond_estimate(vd, jac, wcat, damping)
    out["jac_md1e4_p05"] = vd.VectorSpline2D().jacobian((se, sn), (se, sn))
    np.savez_compressed(os.path.join(GOLDEN, "vector_small.npz"), **out, **ver)

    # ------------------------------------------------------------------
    # 4. SplineCV / cross_val_score on a small problem
    # ------------------------------------------------------------------
    tab = cb.scatter(size=300, random_state=0)
    e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
    dampings = [1e-6, 1e-4, 1e-2, 1.0]
    mindists = [0, 1e-3, 10.0]
    cv = vd.SplineCV(dampings=dampings, mindists=mindists).fit((e, n), d)
    out = dict(east=e, north=n, data=d, dampings=np.array(dampings), mindists=np.array(mindists),
               scores=np.asarray(cv.scores_), best_damping=cv.damping_, best_mindist=cv.mindist_,
               force=cv.force_, grid_east=grid[0], grid_north=grid[1], grid=cv.predict(grid))
    w300 = np.random.RandomState(5).uniform(0.2, 5, e.size)
    out["weights"] = w300
    out["cvs_weighted"] = vd.cross_val_score(vd.Spline(damping=1e-4), (e, n), d, weights=w300)
    out["cvs_neg_rmse"] = vd.cross_val_score(vd.Spline(damping=1e-4), (e, n), d,
                                             scoring="neg_root_mean_squared_error")
    from sklearn.model_selection import KFold

    folds = list(KFold(shuffle=True, random_state=0, n_splits=5).split(np.transpose([e, n])))
    for k, (tr, te) in enumerate(folds):
        out["train_%d" % k], out["test_%d" % k] = tr, te
    np.savez_compressed(os.path.join(GOLDEN, "splinecv_small.npz"), **out, **ver)

    # ------------------------------------------------------------------
    # 5. BASELINE config 1 (5000 pts, damping 1e-8, grid 500x500) summary +
    #    a well-conditioned twin where the 1e-6 bars are decidable
    # ------------------------------------------------------------------
    tab = cb.scatter(size=5000, random_state=0)
    e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
    out = dict(east_head=e[:16], north_head=n[:16], data_head=d[:16])
    full = vd.grid_coordinates((e.min(), e.max(), n.min(), n.max()), shape=(500, 500))
    sub = (full[0][::10, ::10], full[1][::10, ::10])
    for tag, damping, mindist in (("cfg1", 1e-8, None), ("wellcond", 1e-1, None),
                                  ("cfg2like", 1e-6, 1e-5)):
        sp = vd.Spline(damping=damping, mindist=mindist).fit((e, n), d)
        out["force_" + tag] = sp.force_
        out["grid_sub_" + tag] = sp.predict(sub)
        out["pred_data_" + tag] = sp.predict((e[:500], n[:500]))
        out["cond_" + tag] = cond_estimate(vd, sp.jacobian((e, n), (e, n)), None, damping)
    np.savez_compressed(os.path.join(GOLDEN, "cfg1_5k.npz"), **out, **ver)
    for name in sorted(os.listdir(GOLDEN)):
        print(name, os.path.getsize(os.path.join(GOLDEN, name)))


if __name__ == "__main__":
    main()
