"""
CPU oracle for the Verde Green's-function spline hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``verde_b200/`` imports this module.
It may be imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, as the checker and the CPU
yardstick -- never as the product path.

Parity status: PINNED.  ``tests/test_oracle.py`` checks every function here
against (a) golden vectors in ``tests/golden/*.npz`` that were produced by the
UNMODIFIED reference (``oracle/make_golden.py`` imports /root/reference with
numba 0.65.0, scikit-learn 1.9.0, scipy 1.18.1, numpy 2.3.5) and (b) the live
reference when /root/reference is present (build container only).

What is restated, and from where (all paths relative to /root/reference unless
prefixed ``sklearn/`` or ``scipy/``):

=====================  =========================================================
function here          reference
=====================  =========================================================
greens_func            verde/spline.py:564-605 (numpy twin + numba scalar)
jacobian               verde/spline.py:617-650
predict                verde/spline.py:608-639
greens_func_2d         verde/vector.py:393-405
jacobian_2d            verde/vector.py:424-475
predict_2d             verde/vector.py:408-458
column_scale           sklearn/preprocessing/_data.py:83-133,1042-1078 and
                       sklearn/utils/extmath.py:1203-1240 (StandardScaler,
                       with_mean=False: population std about the column mean)
least_squares          verde/base/least_squares.py:17-71 with
                       sklearn/linear_model/_base.py:223-278 (sqrt(w) row
                       rescale), _ridge.py:215-227 (primal Cholesky solve),
                       _ridge.py:230-290 (dual/kernel form when m > n),
                       scipy.linalg.solve(assume_a="pos")
least_squares_sklearn  the literal call sequence of verde/base/least_squares.py
                       against the installed scikit-learn (third-party, NOT
                       vendored in the reference: setup.cfg:48 scikit-learn>=1.0)
spline_fit / *_fit_2d  verde/spline.py:454-464, verde/vector.py:270-289
r2_score               sklearn.metrics.r2_score as used by
                       verde/base/utils.py:50-66 (score_estimator)
kfold_indices          sklearn KFold(shuffle=True, random_state=0, n_splits=5),
                       the default of verde/model_selection.py:759-760
=====================  =========================================================

The C twin of the Green's kernels (oracle/greens_oracle.c, OpenMP) is used when
built; otherwise the numpy forms below are used.  Both are checked against the
same golden vectors.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libverde_oracle.so")
_LIB = None


def build_c_oracle(force=False):
    """Compile oracle/greens_oracle.c with the committed Makefile."""
    src = os.path.join(_HERE, "greens_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CC": "gcc"})
    return _LIB_PATH


def _lib():
    global _LIB
    if _LIB is None:
        build_c_oracle()
        lib = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        sz = ctypes.c_size_t
        dbl = ctypes.c_double
        lib.oracle_greens.restype = dbl
        lib.oracle_greens.argtypes = [dbl, dbl, dbl]
        lib.oracle_jacobian.restype = None
        lib.oracle_jacobian.argtypes = [dp, dp, sz, dp, dp, sz, dbl, dp]
        lib.oracle_predict.restype = None
        lib.oracle_predict.argtypes = [dp, dp, sz, dp, dp, dp, sz, dbl, dp]
        lib.oracle_jacobian_2d.restype = None
        lib.oracle_jacobian_2d.argtypes = [dp, dp, sz, dp, dp, sz, dbl, dbl, dp]
        lib.oracle_predict_2d.restype = None
        lib.oracle_predict_2d.argtypes = [dp, dp, sz, dp, dp, dp, sz, dbl, dbl, dp, dp]
        _LIB = lib
    return _LIB


def _f64(a):
    return np.ascontiguousarray(np.ravel(np.atleast_1d(a)), dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# --------------------------------------------------------------------------
# Scalar biharmonic spline  (verde/spline.py)
# --------------------------------------------------------------------------
def greens_func(east, north, mindist):
    """numpy restatement of verde/spline.py:564-584 (same branch, same formulas)."""
    east = np.asarray(east, dtype=np.float64)
    north = np.asarray(north, dtype=np.float64)
    distance = np.sqrt(east**2 + north**2)
    distance = distance + mindist
    result = np.empty_like(distance)
    small = distance < 1
    big = ~small
    with np.errstate(all="ignore"):
        result[small] = distance[small] * (
            np.log(distance[small] ** distance[small]) - distance[small]
        )
        result[big] = distance[big] ** 2 * (np.log(distance[big]) - 1)
    return result


def jacobian(east, north, force_east, force_north, mindist, engine="c"):
    """(n_data, n_forces) row-major float64 Jacobian; verde/spline.py:617-650."""
    e, n, fe, fn = _f64(east), _f64(north), _f64(force_east), _f64(force_north)
    if engine == "numpy":
        return greens_func(e[:, None] - fe[None, :], n[:, None] - fn[None, :], mindist)
    jac = np.empty((e.size, fe.size), dtype=np.float64)
    _lib().oracle_jacobian(_ptr(e), _ptr(n), e.size, _ptr(fe), _ptr(fn), fe.size,
                           float(mindist), _ptr(jac))
    return jac


def predict(east, north, force_east, force_north, mindist, forces, engine="c"):
    """result[i] = sum_j g(.)*forces[j], j ascending; verde/spline.py:608-639."""
    shape = np.broadcast(np.asarray(east), np.asarray(north)).shape
    e, n, fe, fn, f = _f64(east), _f64(north), _f64(force_east), _f64(force_north), _f64(forces)
    out = np.empty(e.size, dtype=np.float64)
    if engine == "numpy":
        out[:] = 0
        for j in range(f.size):  # verde/spline.py:611-613
            out += greens_func(e - fe[j], n - fn[j], mindist) * f[j]
    else:
        _lib().oracle_predict(_ptr(e), _ptr(n), e.size, _ptr(fe), _ptr(fn), _ptr(f), f.size,
                              float(mindist), _ptr(out))
    return out.reshape(shape)


# --------------------------------------------------------------------------
# 2-component elastic spline  (verde/vector.py)
# --------------------------------------------------------------------------
def greens_func_2d(east, north, mindist, poisson):
    """verde/vector.py:393-405."""
    east = np.asarray(east, dtype=np.float64)
    north = np.asarray(north, dtype=np.float64)
    distance = np.sqrt(east**2 + north**2)
    distance = distance + mindist
    ln_r = (3 - poisson) * np.log(distance)
    over_r2 = (1 + poisson) / distance**2
    green_ee = ln_r + over_r2 * north**2
    green_nn = ln_r + over_r2 * east**2
    green_ne = -over_r2 * east * north
    return green_ee, green_nn, green_ne


def jacobian_2d(east, north, force_east, force_north, mindist, poisson, engine="c"):
    """(2n, 2m) block Jacobian [[ee, ne], [ne, nn]]; verde/vector.py:424-475."""
    e, n, fe, fn = _f64(east), _f64(north), _f64(force_east), _f64(force_north)
    jac = np.empty((2 * e.size, 2 * fe.size), dtype=np.float64)
    if engine == "numpy":
        ee, nn, ne = greens_func_2d(e[:, None] - fe[None, :], n[:, None] - fn[None, :],
                                    mindist, poisson)
        npts, nf = e.size, fe.size
        jac[:npts, :nf] = ee
        jac[npts:, nf:] = nn
        jac[:npts, nf:] = ne
        jac[npts:, :nf] = ne
    else:
        _lib().oracle_jacobian_2d(_ptr(e), _ptr(n), e.size, _ptr(fe), _ptr(fn), fe.size,
                                  float(mindist), float(poisson), _ptr(jac))
    return jac


def predict_2d(east, north, force_east, force_north, mindist, poisson, forces, engine="c"):
    """(vec_east, vec_north); forces stacked east-then-north; verde/vector.py:408-458."""
    shape = np.broadcast(np.asarray(east), np.asarray(north)).shape
    e, n, fe, fn, f = _f64(east), _f64(north), _f64(force_east), _f64(force_north), _f64(forces)
    m = f.size // 2
    ve = np.empty(e.size, dtype=np.float64)
    vn = np.empty(e.size, dtype=np.float64)
    if engine == "numpy":
        ve[:] = 0
        vn[:] = 0
        for j in range(m):
            ee, nn, ne = greens_func_2d(e - fe[j], n - fn[j], mindist, poisson)
            ve += ee * f[j] + ne * f[j + m]
            vn += ne * f[j] + nn * f[j + m]
    else:
        _lib().oracle_predict_2d(_ptr(e), _ptr(n), e.size, _ptr(fe), _ptr(fn), _ptr(f), m,
                                 float(mindist), float(poisson), _ptr(ve), _ptr(vn))
    return ve.reshape(shape), vn.reshape(shape)


# --------------------------------------------------------------------------
# least squares  (verde/base/least_squares.py + scikit-learn + scipy)
# --------------------------------------------------------------------------
def column_scale(jac):
    """
    ``StandardScaler(with_mean=False).fit(jac).scale_``.

    Population standard deviation ABOUT THE COLUMN MEAN (with_mean=False only
    skips the centring of the transformed data, not of the variance), computed
    with the corrected two-pass formula of sklearn/utils/extmath.py:1203-1240
    and the constant-column rule of sklearn/preprocessing/_data.py:83-133.
    """
    jac = np.asarray(jac, dtype=np.float64)
    n = jac.shape[0]
    new_sum = np.sum(jac, axis=0)
    mean = new_sum / n
    temp = jac - mean
    correction = np.sum(temp, axis=0)
    temp **= 2
    unnorm = np.sum(temp, axis=0)
    unnorm -= correction**2 / n
    var = unnorm / n
    # _is_constant_feature (sklearn/preprocessing/_data.py:83-98)
    eps = np.finfo(np.float64).eps
    upper_bound = n * eps * var + (n * mean * eps) ** 2
    constant = var <= upper_bound
    var = np.where(constant, 0.0, var)
    scale = np.sqrt(var)
    # _handle_zeros_in_scale (sklearn/preprocessing/_data.py:101-133)
    scale = np.where(constant | (scale < 10 * eps), 1.0, scale)
    return scale


# scipy.linalg.lstsq's cond in the damping=None branch = LinearRegression.tol's default in scikit-learn 1.9.0,
# the version the goldens were made with (tests/golden/damping_none.npz stores it as ``lstsq_cond``)
LSTSQ_COND = 1e-6


def least_squares(jac, data, weights, damping):
    """
    Explicit-algebra restatement of ``verde.base.least_squares``: the damped branch below, and
    ``damping=None`` as a truncated SVD (see the comment in the body).

    ``Js = jac / s``;  rows rescaled by sqrt(w);  primal
    ``(Js^T Js + damping I) c = Js^T d`` solved with
    ``scipy.linalg.solve(assume_a="pos")`` when n >= m, the dual
    ``(Js Js^T + damping I) a = d, c = Js^T a`` when m > n
    (sklearn/linear_model/_ridge.py:746-761);  returns ``c / s``.
    """
    import scipy.linalg

    jac = np.array(jac, dtype=np.float64, copy=True)
    d = np.array(np.ravel(data), dtype=np.float64, copy=True)
    n, m = jac.shape
    scale = column_scale(jac)
    jac /= scale
    if weights is not None:
        sw = np.sqrt(np.asarray(np.ravel(weights), dtype=np.float64))
        jac *= sw[:, None]
        d *= sw
    if damping is None:
        # verde/base/least_squares.py:65-66 -> LinearRegression(fit_intercept=False) ->
        # scipy.linalg.lstsq(X, y, cond=LinearRegression.tol) (sklearn/linear_model/_base.py, scikit-learn 1.9):
        # LAPACK gelsd = minimum-norm solution over the singular values sigma > cond * sigma_max
        u, sv, vt = np.linalg.svd(jac, full_matrices=False)
        keep = sv > LSTSQ_COND * sv[0]
        coef = vt[keep].T @ ((u[:, keep].T @ d) / sv[keep])
        return coef / scale
    if m > n:
        kernel = jac @ jac.T
        kernel.flat[:: n + 1] += damping
        dual = scipy.linalg.solve(kernel, d, assume_a="pos", overwrite_a=True)
        coef = jac.T @ dual
    else:
        gram = jac.T @ jac
        rhs = jac.T @ d
        gram.flat[:: m + 1] += damping
        coef = scipy.linalg.solve(gram, rhs, assume_a="pos", overwrite_a=True)
    return coef / scale


def least_squares_sklearn(jac, data, weights, damping):
    """The literal call sequence of verde/base/least_squares.py:63-70."""
    from sklearn.linear_model import LinearRegression, Ridge
    from sklearn.preprocessing import StandardScaler

    scaler = StandardScaler(copy=True, with_mean=False, with_std=True)
    jac = scaler.fit_transform(np.asarray(jac, dtype=np.float64))
    regr = LinearRegression(fit_intercept=False) if damping is None else Ridge(alpha=damping, fit_intercept=False)
    regr.fit(jac, np.ravel(data), sample_weight=weights)
    return regr.coef_ / scaler.scale_


def normal_equations(jac, data, weights):
    """
    The UNSCALED quantities the fused GPU Gram kernel accumulates (SURVEY.md
    section 7.2): G = J^T W J, b = J^T W d, c1 = J^T 1, c2 = diag(J^T J).
    """
    jac = np.asarray(jac, dtype=np.float64)
    d = np.asarray(np.ravel(data), dtype=np.float64)
    if weights is None:
        jw = jac
    else:
        jw = jac * np.asarray(np.ravel(weights), dtype=np.float64)[:, None]
    gram = jw.T @ jac
    rhs = jw.T @ d
    c1 = jac.sum(axis=0)
    c2 = np.einsum("ij,ij->j", jac, jac)
    return gram, rhs, c1, c2


def spline_fit(east, north, data, weights=None, damping=None, mindist=0.0,
               force_east=None, force_north=None):
    """verde/spline.py:454-464 -> force_ (forces at the data when force_* is None)."""
    e, n = _f64(east), _f64(north)
    fe = e.copy() if force_east is None else _f64(force_east)
    fn = n.copy() if force_north is None else _f64(force_north)
    jac = jacobian(e, n, fe, fn, mindist)
    return least_squares(jac, data, weights, damping)


def vector_fit_2d(east, north, data_east, data_north, weights=None, damping=None,
                  mindist=10e3, poisson=0.5, force_east=None, force_north=None):
    """verde/vector.py:270-289 -> force_ (length 2m, east forces then north forces)."""
    e, n = _f64(east), _f64(north)
    fe = e.copy() if force_east is None else _f64(force_east)
    fn = n.copy() if force_north is None else _f64(force_north)
    jac = jacobian_2d(e, n, fe, fn, mindist, poisson)
    data = np.concatenate([_f64(data_east), _f64(data_north)])
    if weights is not None:
        weights = np.concatenate([_f64(w) for w in weights])
    return least_squares(jac, data, weights, damping)


# --------------------------------------------------------------------------
# scoring / CV helpers
# --------------------------------------------------------------------------
def r2_score(y_true, y_pred, sample_weight=None):
    """sklearn.metrics.r2_score for 1-D targets (used via verde/base/utils.py:50-66)."""
    y_true = np.ravel(np.asarray(y_true, dtype=np.float64))
    y_pred = np.ravel(np.asarray(y_pred, dtype=np.float64))
    w = np.ones_like(y_true) if sample_weight is None else np.ravel(sample_weight)
    numerator = np.sum(w * (y_true - y_pred) ** 2)
    denominator = np.sum(w * (y_true - np.average(y_true, weights=w)) ** 2)
    if denominator == 0:
        return 1.0 if numerator == 0 else 0.0
    return 1.0 - numerator / denominator


def kfold_indices(n_samples, n_splits=5, random_state=0):
    """Index sets of KFold(shuffle=True, random_state=0, n_splits=5).split(X)."""
    rng = np.random.RandomState(random_state)
    indices = np.arange(n_samples)
    rng.shuffle(indices)
    fold_sizes = np.full(n_splits, n_samples // n_splits, dtype=int)
    fold_sizes[: n_samples % n_splits] += 1
    current = 0
    folds = []
    for size in fold_sizes:
        test_mask = np.zeros(n_samples, dtype=bool)
        test_mask[indices[current:current + size]] = True
        folds.append((np.flatnonzero(~test_mask), np.flatnonzero(test_mask)))
        current += size
    return folds


def checkerboard(east, north, amplitude=1000.0, region=(0, 5000, -5000, 0)):
    """verde/synthetic.py:112-118 with the default half-extent wavelengths (:65-86)."""
    w_east = (region[1] - region[0]) / 2
    w_north = (region[3] - region[2]) / 2
    return (
        amplitude
        * np.sin((2 * np.pi / w_east) * np.asarray(east))
        * np.cos((2 * np.pi / w_north) * np.asarray(north))
    )
