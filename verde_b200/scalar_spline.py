# Host glue derived from fatiando/verde (BSD 3-Clause): it restates the reference's estimator contract, error and
# warning texts so that this path is a drop-in.
#   Copyright (c) 2017 The Verde Developers.  Distributed under the terms of the BSD 3-Clause License.
#   SPDX-License-Identifier: BSD-3-Clause
#   This code is part of the Fatiando a Terra project (https://www.fatiando.org); see LICENSE-verde.txt.
"""
``Spline`` and ``SplineCV``: drop-ins for ``verde.Spline`` / ``verde.SplineCV``
(verde/spline.py:29-538) whose numerics run on a B200.

Constructor signatures, fitted attributes, warnings and error types follow the
reference; what changes is underneath ``fit`` / ``predict`` / ``jacobian``:

* ``fit``      one fused GPU call (Jacobian panels generated on the fly ->
               DMMA Gram accumulation -> StandardScaler algebra -> damping ->
               blocked Cholesky -> substitution); the n x m Jacobian the reference
               materialises (verde/spline.py:462) never exists.
* ``predict``  matrix-free N-body kernel (verde/spline.py:486-499).
* ``SplineCV`` builds each (fold, mindist) Gram matrix ONCE and re-solves it for
               every damping (the reference refits all 6 x 4 x 5 combinations from
               scratch, verde/spline.py:243-255); jobs are dealt to the ranks of a
               ``torch.distributed`` group when one is initialised.

``damping=None`` (the reference's default) is scikit-learn's ``LinearRegression``
there: the minimum-norm least-squares solution with singular values below
``1e-6 * sigma_max`` dropped (scikit-learn >= 1.9).  Here it is a one-sided Jacobi
SVD of the scaled Jacobian on the device (``vb_svd.cu``, ``kernels.LSTSQ_COND``).
"""
import itertools
import time
import warnings

import numpy as np
from sklearn.base import clone
from sklearn.model_selection import KFold
from sklearn.utils.validation import check_is_fitted

from . import dist, kernels
from .base import BaseGridder, check_fit_input, n_1d_arrays, score_estimator
from .coords import get_region


def warn_weighted_exact_solution(spline, weights):
    "Same UserWarning as verde/spline.py:541-561."
    if weights is not None and spline.damping is None:
        warnings.warn(
            "Weights might have no effect if no regularization is used. "
            "Use damping or specify force positions that are different from the data.",
            stacklevel=2,
        )


def parse_engine(engine):
    """
    The reference accepts 'auto' | 'numba' | 'numpy' (verde/utils.py:64-87).  There
    is exactly one engine here (CUDA); the argument is validated for signature and
    ``clone`` compatibility and otherwise ignored.
    """
    engines = {"auto", "numba", "numpy"}
    if engine not in engines:
        raise ValueError("Invalid engine '{}'. Must be in {}.".format(engine, engines))
    return "b200"


class Spline(BaseGridder):
    """
    Biharmonic spline interpolation with Green's functions [Sandwell1987], on GPU.

    Parameters
    ----------
    mindist : float or None
        Deprecated fudge factor added to every data-force distance (kept, and
        still applied, exactly like the reference).
    damping : float or None
        Positive Tikhonov damping of the column-scaled system (normal equations + Cholesky), or
        ``None`` for the un-regularised truncated-SVD solution (see the module docstring).
    force_coords : None or tuple of arrays
        Easting and northing of the point forces; data locations when None.
    engine : str
        Deprecated, accepted for compatibility.

    Attributes
    ----------
    force_ : array (m,)
    force_coords_ : tuple of two arrays
    region_ : tuple (W, E, S, N)
    fit_info_ : dict
        Device timings, launch counts and ``diag_min/diag_max`` of the Cholesky
        factor (a conditioning indicator) of the last fit.
    """

    def __init__(self, mindist=None, damping=None, force_coords=None, engine="auto"):
        super().__init__()
        self.damping = damping
        self.force_coords = force_coords
        self.engine = engine
        if engine != "auto":
            warnings.warn(
                "The 'engine' parameter of 'verde.Spline' is "
                "deprecated and will be removed in Verde 2.0.0. "
                "The faster and memory efficient numba engine will be "
                "the only option.",
                FutureWarning,
                stacklevel=2,
            )
        if mindist is None:
            self.mindist = 0
        else:
            self.mindist = mindist
            warnings.warn(
                "The mindist parameter of verde.Spline is no longer required "
                "and will be removed in Verde 2.0.0. Use the default value "
                "to obtain the future behavior.",
                FutureWarning,
                stacklevel=2,
            )

    def fit(self, coordinates, data, weights=None):
        "Estimate the forces that fit *data* (same contract as verde/spline.py:426-464)."
        coordinates, data, weights = check_fit_input(coordinates, data, weights)
        warn_weighted_exact_solution(self, weights)
        parse_engine(self.engine)
        self.region_ = get_region(coordinates[:2])
        if self.force_coords is None:
            self.force_coords_ = tuple(i.copy() for i in n_1d_arrays(coordinates, n=2))
        else:
            self.force_coords_ = self.force_coords
        east, north = n_1d_arrays(coordinates, n=2)
        force_east, force_north = n_1d_arrays(self.force_coords_, n=2)
        if dist.sharding_active():
            self.force_ = kernels.fit_sharded(0, east, north, data, weights, force_east, force_north,
                                              self.mindist, 0.0, self.damping)
        else:
            self.force_ = kernels.fit_spline(east, north, data, weights, force_east, force_north,
                                             self.mindist, self.damping)
        self.fit_info_ = dict(kernels.LAST_FIT_INFO)
        return self

    def predict(self, coordinates):
        "Evaluate the spline at *coordinates*; output shape/dtype as verde/spline.py:486-499."
        check_is_fitted(self, ["force_"])
        shape = np.broadcast(*coordinates[:2]).shape
        force_east, force_north = n_1d_arrays(self.force_coords_, n=2)
        east, north = n_1d_arrays(coordinates, n=2)
        if east.size != north.size:
            east, north = (np.ravel(c) for c in np.broadcast_arrays(*coordinates[:2]))
        if dist.sharding_active():
            data = kernels.predict_sharded(east, north, force_east, force_north, self.mindist, self.force_)
            data = data.astype(east.dtype if np.issubdtype(east.dtype, np.floating) else np.float64, copy=False)
        else:
            data = np.empty(east.size, dtype=east.dtype if np.issubdtype(east.dtype, np.floating) else np.float64)
            data = kernels.predict_b200(east, north, force_east, force_north, self.mindist, self.force_, data)
        return data.reshape(shape)

    def jacobian(self, coordinates, force_coords, dtype="float64"):
        "(n_data, n_forces) Jacobian, built on the GPU and returned as a host array."
        force_east, force_north = n_1d_arrays(force_coords, n=2)
        east, north = n_1d_arrays(coordinates, n=2)
        jac = np.empty((east.size, force_east.size), dtype=dtype)
        return kernels.jacobian_b200(east, north, force_east, force_north, self.mindist, jac)


def select(arrays, index):
    "Index every array of a tuple (verde/model_selection.py:791-821); None tuples pass through."
    if arrays is None or any(i is None for i in arrays):
        return arrays
    return tuple(np.ravel(i)[index] for i in arrays)


def fit_score(estimator, train_data, test_data, scoring):
    "verde/model_selection.py:779-788."
    estimator.fit(*train_data)
    if scoring is None:
        return estimator.score(*test_data)
    return score_estimator(scoring, estimator, *test_data)


def cross_val_score(estimator, coordinates, data, weights=None, cv=None, client=None, delayed=False,
                    scoring=None):
    """
    Score an estimator by cross-validation (verde/model_selection.py:582-776).

    Default splitter: ``KFold(shuffle=True, random_state=0, n_splits=5)``.  Folds
    run one after the other on the current GPU (each fit already fills it);
    ``delayed`` / ``client`` (dask) are not part of this path.
    """
    if client is not None:
        warnings.warn(
            "The 'client' parameter of 'verde.cross_val_score' is deprecated "
            "and will be removed in Verde 2.0.0. "
            "Use the 'delayed' parameter instead.",
            FutureWarning,
            stacklevel=2,
        )
    if delayed:
        raise NotImplementedError("delayed=True needs dask, which is outside the B200 path")
    coordinates, data, weights = check_fit_input(coordinates, data, weights, unpack=False)
    if cv is None:
        cv = KFold(shuffle=True, random_state=0, n_splits=5)
    feature_matrix = np.transpose(n_1d_arrays(coordinates, 2))
    fit_args = (coordinates, data, weights)
    scores = []
    with warnings.catch_warnings():
        warnings.filterwarnings("ignore", message="The default scoring will change", category=FutureWarning)
        for train_index, test_index in cv.split(feature_matrix):
            scores.append(
                fit_score(
                    clone(estimator),
                    tuple(select(i, train_index) for i in fit_args),
                    tuple(select(i, test_index) for i in fit_args),
                    scoring,
                )
            )
    return np.asarray(scores)


class _FrozenSpline:
    "Minimal predictor handed to the scorer inside SplineCV's sweep."

    def __init__(self, force, force_coords, mindist):
        self.force_, self.force_coords_, self.mindist = force, force_coords, mindist

    def predict(self, coordinates):
        east, north = n_1d_arrays(coordinates, n=2)
        out = np.empty(east.size, dtype=np.float64)
        fe, fn = n_1d_arrays(self.force_coords_, n=2)
        return kernels.predict_b200(east, north, fe, fn, self.mindist, self.force_, out)


class SplineCV(BaseGridder):
    """
    Cross-validated biharmonic spline (verde/spline.py:29-312).

    Scores every ``(mindist, damping)`` of ``itertools.product(mindists, dampings)``
    (mindist-major, as in the reference) with the mean cross-validation score,
    keeps the best and refits it on all data.  ``scores_`` has one entry per
    combination in that order.

    The sweep is restructured for the GPU without changing its result: the
    Jacobian, and therefore the Gram matrix, depends only on (fold, mindist);
    damping only shifts the diagonal of the scaled system.  Each Gram matrix is
    built once, kept in HBM and factored once per damping.
    """

    def __init__(self, mindists=None, dampings=(1e-10, 1e-5, 1e-1), force_coords=None, engine="auto",
                 cv=None, client=None, delayed=False, scoring=None):
        super().__init__()
        self.dampings = dampings
        if mindists is None:
            self.mindists = [0]
        else:
            self.mindists = mindists
            warnings.warn(
                "The mindists parameter of verde.SplineCV is no longer "
                "required and will be removed in Verde 2.0.0. Use the default "
                "value to obtain the future behavior.",
                FutureWarning,
                stacklevel=2,
            )
        self.force_coords = force_coords
        self.engine = engine
        self.cv = cv
        self.client = client
        self.delayed = delayed
        self.scoring = scoring
        if engine != "auto":
            warnings.warn(
                "The 'engine' parameter of 'verde.SplineCV' is "
                "deprecated and will be removed in Verde 2.0.0. "
                "The faster and memory efficient numba engine will be "
                "the only option.",
                FutureWarning,
                stacklevel=2,
            )
        if client is not None:
            warnings.warn(
                "The 'client' parameter of 'verde.SplineCV' is "
                "deprecated and will be removed in Verde 2.0.0. "
                "Use the 'delayed' parameter instead.",
                FutureWarning,
                stacklevel=2,
            )

    def fit(self, coordinates, data, weights=None):
        "Tune (mindist, damping) by cross-validation, then fit the best on all data."
        if self.delayed:
            raise NotImplementedError("delayed=True needs dask, which is outside the B200 path")
        parse_engine(self.engine)
        combos = list(itertools.product(self.mindists, self.dampings))
        coords_c, data_c, weights_c = check_fit_input(coordinates, data, weights, unpack=False)
        cv = self.cv if self.cv is not None else KFold(shuffle=True, random_state=0, n_splits=5)
        feature_matrix = np.transpose(n_1d_arrays(coords_c, 2))
        folds = list(cv.split(feature_matrix))
        fit_args = (coords_c, data_c, weights_c)
        scoring = "r2" if self.scoring is None else self.scoring

        mindists, dampings = list(self.mindists), list(self.dampings)
        table = np.zeros((len(mindists), len(dampings), len(folds)))
        jobs = [(f, i) for f in range(len(folds)) for i in range(len(mindists))]
        sweep = {"jobs": 0, "gram_ms": 0.0, "factor_ms": 0.0, "solve_ms": 0.0, "gemm_ms": 0.0, "gemm_flops": 0.0,
                 "score_s": 0.0, "split_s": 0.0, "accumulate_s": 0.0, "solve_s": 0.0}
        # A failure on one rank (a non-positive-definite system, out of memory) must not strand the others in the
        # all-reduce of the score table: the rank stops working, still joins the collective, and every rank raises.
        failure = None
        for job in dist.round_robin(len(jobs)):
            fold, imd = jobs[job]
            t_job = time.perf_counter()
            train_index, test_index = folds[fold]
            train = tuple(select(a, train_index) for a in fit_args)
            test = tuple(select(a, test_index) for a in fit_args)
            east, north = n_1d_arrays(train[0], n=2)
            if self.force_coords is None:
                force_coords = (east.copy(), north.copy())
            else:
                force_coords = n_1d_arrays(self.force_coords, n=2)
            train_w = None if train[2][0] is None else train[2][0]
            kernels._warn_underdetermined(east.size, force_coords[0].size)
            sweep["split_s"] += time.perf_counter() - t_job
            try:
                t0 = time.perf_counter()
                system = kernels.GramSystem(force_coords[0].size, peer_exchange=False)  # a rank-local job
                system.accumulate(0, east, north, train[1][0], train_w, force_coords[0], force_coords[1],
                                  mindists[imd])
                sweep["accumulate_s"] += time.perf_counter() - t0
                t0 = time.perf_counter()
                # all dampings of this (fold, mindist) are factored and solved side by side on the device;
                # a None among them is the truncated-SVD solve, which shares nothing with the Gram matrix
                damped = [k for k, damp in enumerate(dampings) if damp is not None]
                forces = np.empty((len(dampings), force_coords[0].size))
                if damped:
                    forces[damped] = system.solve_multi(east.size, [dampings[k] for k in damped])
                for k, damp in enumerate(dampings):
                    if damp is None:
                        forces[k] = kernels.fit_spline(east, north, train[1][0], train_w, force_coords[0],
                                                       force_coords[1], mindists[imd], None)
                sweep["solve_s"] += time.perf_counter() - t0
            except (np.linalg.LinAlgError, MemoryError, kernels._cabi.EngineError) as exc:
                failure = exc
                table[imd, :, fold] = np.nan
                break
            for key in ("gram_ms", "factor_ms", "solve_ms", "gemm_ms", "gemm_flops"):
                sweep[key] += getattr(system.info, key)
            t0 = time.perf_counter()
            for idamp in range(len(dampings)):
                frozen = _FrozenSpline(forces[idamp], force_coords, mindists[imd])
                table[imd, idamp, fold] = score_estimator(scoring, frozen, *test)
            sweep["score_s"] += time.perf_counter() - t0
            sweep["jobs"] += 1
            del system
        self.sweep_info_ = sweep
        table = dist.allreduce_sum_numpy(table)
        lowest, _ = dist.agree_on_status(-1 if failure is not None else 0)
        if failure is not None:
            raise failure
        if lowest < 0:
            raise np.linalg.LinAlgError("SplineCV: a (fold, mindist) job failed on another rank")
        scores = table.mean(axis=2).reshape(-1)
        best = int(np.argmax(scores))
        self.spline_ = Spline(mindist=None, damping=combos[best][1], force_coords=self.force_coords,
                              engine="auto")
        # bypass the FutureWarning of a non-None mindist but keep the reference's value
        self.spline_.mindist = combos[best][0]
        self.spline_.engine = self.engine
        self.spline_.fit(coordinates, data, weights=weights)
        self.scores_ = scores
        return self

    @property
    def force_(self):
        "The estimated forces that fit the data."
        return self.spline_.force_

    @property
    def region_(self):
        "The bounding region of the data used to fit the spline"
        return self.spline_.region_

    @property
    def damping_(self):
        "The optimal damping parameter"
        return self.spline_.damping

    @property
    def mindist_(self):
        "The optimal mindist parameter"
        return self.spline_.mindist

    @property
    def force_coords_(self):
        "The optimal force locations"
        return self.spline_.force_coords_

    def predict(self, coordinates):
        "Evaluate the best spline at *coordinates*."
        check_is_fitted(self, ["spline_"])
        return self.spline_.predict(coordinates)
