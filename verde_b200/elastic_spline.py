# Host glue derived from fatiando/verde (BSD 3-Clause): it restates the reference's estimator contract, error and
# warning texts so that this path is a drop-in.
#   Copyright (c) 2017 The Verde Developers.  Distributed under the terms of the BSD 3-Clause License.
#   SPDX-License-Identifier: BSD-3-Clause
#   This code is part of the Fatiando a Terra project (https://www.fatiando.org); see LICENSE-verde.txt.
"""
``VectorSpline2D``: drop-in for ``verde.VectorSpline2D`` (verde/vector.py:144-390),
the elastically coupled 2-component spline of [SandwellWessel2016], on a B200.

The 2n x 2m block system ``[[J_ee, J_ne], [J_ne, J_nn]] [f_e; f_n] = [d_e; d_n]``
is never formed: the GPU generates its row panels on the fly and accumulates the
2m x 2m normal equations (see scalar_spline.py / DESIGN.md).  Quirks of the
reference that callers may rely on are kept: data and weights are tuples of two
arrays stacked east-then-north (vector.py:279-284); with ``force_coords=None``
the data locations are written into the constructor parameter ``force_coords``
on first fit and reused afterwards (vector.py:285-286); ``force_`` holds the m
east forces followed by the m north forces (vector.py:456-457).
"""
import warnings

import numpy as np
from sklearn.utils.validation import check_is_fitted

from . import dist, kernels
from .base import BaseGridder, check_fit_input, n_1d_arrays
from .coords import get_region
from .scalar_spline import parse_engine, warn_weighted_exact_solution


class VectorSpline2D(BaseGridder):
    """
    Elastically coupled interpolation of 2-component vector data.

    Parameters
    ----------
    poisson : float
        Poisson's ratio of the elastic sheet (couples the two components).
    mindist : float
        Added to every data-force distance (default 10e3, as the reference).
    damping : float
        Positive damping of the column-scaled system, or ``None`` (the reference default) for the
        truncated-SVD minimum-norm solution (``kernels.LSTSQ_COND``).
    force_coords : None or tuple of arrays
    engine : str
        Deprecated, accepted for compatibility.

    Attributes
    ----------
    force_ : array (2m,)
    region_ : tuple (W, E, S, N)
    """

    def __init__(self, poisson=0.5, mindist=10e3, damping=None, force_coords=None, engine="auto"):
        super().__init__()
        self.poisson = poisson
        self.mindist = mindist
        self.damping = damping
        self.force_coords = force_coords
        self.engine = engine
        if engine != "auto":
            warnings.warn(
                "The 'engine' parameter of 'verde.VectorSpline2D' is "
                "deprecated and will be removed in Verde 2.0.0. "
                "The faster and memory efficient numba engine will be "
                "the only option.",
                FutureWarning,
                stacklevel=2,
            )

    def fit(self, coordinates, data, weights=None):
        "Estimate the 2m forces from ``data = (east_component, north_component)``."
        coordinates, data, weights = check_fit_input(coordinates, data, weights, unpack=False)
        if len(data) != 2:
            raise ValueError("Need two data components. Only {} given.".format(len(data)))
        parse_engine(self.engine)
        self.region_ = get_region(coordinates[:2])
        if any(w is not None for w in weights):
            weights = np.concatenate([np.ravel(w) for w in weights])
        else:
            weights = None
        warn_weighted_exact_solution(self, weights)
        if self.force_coords is None:
            self.force_coords = tuple(i.copy() for i in n_1d_arrays(coordinates, n=2))
        east, north = n_1d_arrays(coordinates, n=2)
        force_east, force_north = n_1d_arrays(self.force_coords, n=2)
        if dist.sharding_active():
            comps = tuple(np.ravel(c) for c in data)
            wcomps = None if weights is None else (weights[: east.size], weights[east.size:])
            self.force_ = kernels.fit_sharded(1, east, north, comps, wcomps, force_east, force_north,
                                              self.mindist, self.poisson, self.damping)
        else:
            stacked = np.concatenate([np.ravel(c) for c in data])
            self.force_ = kernels.fit_spline_2d(east, north, stacked, weights, force_east, force_north,
                                                self.mindist, self.poisson, self.damping)
        self.fit_info_ = dict(kernels.LAST_FIT_INFO)
        return self

    def predict(self, coordinates):
        "``(east_component, north_component)`` at *coordinates* (verde/vector.py:291-345)."
        check_is_fitted(self, ["force_"])
        force_east, force_north = n_1d_arrays(self.force_coords, n=2)
        east, north = n_1d_arrays(coordinates, n=2)
        cast = np.broadcast(*coordinates[:2])
        dtype = east.dtype if np.issubdtype(east.dtype, np.floating) else np.float64
        if dist.sharding_active():
            comps = kernels.predict_sharded(east, north, force_east, force_north, self.mindist,
                                            self.force_, poisson=self.poisson)
            comps = tuple(c.astype(dtype, copy=False) for c in comps)
        else:
            comps = (np.empty(cast.size, dtype=dtype), np.empty(cast.size, dtype=dtype))
            comps = kernels.predict_2d_b200(east, north, force_east, force_north, self.mindist,
                                            self.poisson, self.force_, comps[0], comps[1])
        return tuple(c.reshape(cast.shape) for c in comps)

    def jacobian(self, coordinates, force_coords, dtype="float64"):
        "(2 n_data, 2 n_forces) block Jacobian, built on the GPU, returned on the host."
        force_east, force_north = n_1d_arrays(force_coords, n=2)
        east, north = n_1d_arrays(coordinates, n=2)
        jac = np.empty((east.size * 2, force_east.size * 2), dtype=dtype)
        return kernels.jacobian_2d_b200(east, north, force_east, force_north, self.mindist,
                                        self.poisson, jac)
