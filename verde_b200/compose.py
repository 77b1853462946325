# Host glue derived from fatiando/verde (BSD 3-Clause): it restates the reference's estimator contract, error and
# warning texts so that this path is a drop-in.
#   Copyright (c) 2017 The Verde Developers.  Distributed under the terms of the BSD 3-Clause License.
#   SPDX-License-Identifier: BSD-3-Clause
#   This code is part of the Fatiando a Terra project (https://www.fatiando.org); see LICENSE-verde.txt.
"""
Composition wrappers that every Verde example puts around the spline:
``Vector`` (one scalar gridder per data component, verde/vector.py:28-141) and
``Chain`` (gridders applied in sequence to each other's residuals,
verde/chain.py:17-138).  Pure host glue: they only ever call ``fit`` /
``filter`` / ``predict`` on their members, so the B200 gridders slot in
unchanged.
"""
import numpy as np
from sklearn.utils.validation import check_is_fitted

from . import dist, kernels
from .base import BaseGridder, check_data, check_fit_input, n_1d_arrays
from .coords import get_region
from .scalar_spline import Spline, parse_engine


class Vector(BaseGridder):
    """
    Fit one estimator per component of multi-component data.

    ``components`` is a list of gridders (e.g. ``[Spline(damping=...)] * 2`` must be
    two distinct instances, as in Verde); ``predict`` returns one array per component.
    """

    def __init__(self, components):
        super().__init__()
        self.components = components

    def fit(self, coordinates, data, weights=None):
        if not isinstance(data, tuple):
            raise ValueError("Data must be a tuple of arrays. {} given.".format(type(data)))
        if weights is not None and not isinstance(weights, tuple):
            raise ValueError("Weights must be a tuple of arrays. {} given.".format(type(weights)))
        coordinates, data, weights = check_fit_input(coordinates, data, weights)
        self.region_ = get_region(coordinates[:2])
        if self._splines_share_a_system(data, weights):
            self._fit_shared(coordinates, data, weights)
        else:
            for estimator, data_comp, weight_comp in zip(self.components, data, weights):
                estimator.fit(coordinates, data_comp, weight_comp)
        return self

    def _splines_share_a_system(self, data, weights):
        """
        True when every component is a damped :class:`Spline` with the same parameters and the
        same (or no) weights: then all components have the SAME Jacobian and normal matrix and only
        the right-hand sides differ, so one Gram accumulation + one Cholesky serves them all.
        """
        comps = self.components
        if not isinstance(data, tuple) or len(comps) < 2 or len(comps) != len(data):
            return False
        first = comps[0]
        if any(type(c) is not Spline for c in comps) or first.damping is None:
            return False
        if any(c.damping != first.damping or c.mindist != first.mindist or c.engine != first.engine
               or c.force_coords is not first.force_coords for c in comps[1:]):
            return False
        if all(w is None for w in weights):
            return True
        return all(w is not None and np.array_equal(w, weights[0]) for w in weights)

    def _fit_shared(self, coordinates, data, weights):
        first = self.components[0]
        parse_engine(first.engine)
        east, north = n_1d_arrays(coordinates, n=2)
        if first.force_coords is None:
            force_coords = tuple(i.copy() for i in (east, north))
        else:
            force_coords = first.force_coords
        force_east, force_north = n_1d_arrays(force_coords, n=2)
        w = weights[0]
        kernels._warn_underdetermined(east.size, force_east.size)
        east, north = (np.ascontiguousarray(c, dtype=np.float64) for c in (east, north))
        # Under sharding (one process per GPU) a rank accumulates its contiguous row range, the Gram matrix and the
        # right-hand sides are all-reduced, and every rank factors the (identical) reduced matrix once and solves
        # every component against it: all ranks end with the same forces.
        sharded = dist.sharding_active()
        a, b = dist.shard_bounds(east.size) if sharded else (0, east.size)
        w_rows = None if w is None else np.ravel(w)[a:b]
        system, failure = None, None
        try:
            system = kernels.GramSystem(force_east.size, peer_exchange=False)
            system.accumulate(0, east[a:b], north[a:b], np.ravel(data[0])[a:b], w_rows, force_east, force_north,
                              first.mindist)
        except (MemoryError, ValueError, kernels._cabi.EngineError) as exc:
            failure = exc
        if sharded:
            lowest, _ = dist.agree_on_status(-1 if failure is not None else 0)
            if failure is None and lowest < 0:
                failure = kernels._cabi.EngineError("Vector: the shared Gram accumulation failed on another rank")
        if failure is not None:
            raise failure
        if sharded:
            system.allreduce()
        system.factor(east.size, first.damping)
        region = get_region(coordinates[:2])
        for k, (estimator, data_comp) in enumerate(zip(self.components, data)):
            if k == 0:
                rhs = system.rhs()
            else:
                v = np.ravel(data_comp)[a:b] if w is None else w_rows * np.ravel(data_comp)[a:b]
                rhs = kernels.jacobian_transpose_apply(east[a:b], north[a:b], force_east, force_north, first.mindist, v)
                if sharded:
                    rhs = dist.allreduce_sum_numpy(rhs)
            estimator.force_ = system.solve_rhs(rhs)
            estimator.region_ = region
            estimator.force_coords_ = (force_coords if first.force_coords is not None
                                       else tuple(i.copy() for i in force_coords))
            estimator.fit_info_ = dict(kernels.LAST_FIT_INFO)

    def _splines_share_forces(self):
        "Fitted Splines on the same force positions and mindist: one Green's evaluation serves every component."
        comps = self.components
        if len(comps) < 2 or any(type(c) is not Spline for c in comps):
            return False
        if any(not hasattr(c, "force_") or c.mindist != comps[0].mindist for c in comps):
            return False
        first = n_1d_arrays(comps[0].force_coords_, n=2)
        return all(c.force_.size == comps[0].force_.size
                   and all(np.array_equal(a, b) for a, b in zip(n_1d_arrays(c.force_coords_, n=2), first))
                   for c in comps[1:])

    def predict(self, coordinates):
        check_is_fitted(self, ["region_"])
        if not self._splines_share_forces():
            return tuple(comp.predict(coordinates) for comp in self.components)
        # same contract as Spline.predict for each component (verde/spline.py:486-499), one pass over the pairs
        shape = np.broadcast(*coordinates[:2]).shape
        east, north = n_1d_arrays(coordinates, n=2)
        if east.size != north.size:
            east, north = (np.ravel(c) for c in np.broadcast_arrays(*coordinates[:2]))
        force_east, force_north = n_1d_arrays(self.components[0].force_coords_, n=2)
        forces = np.stack([np.asarray(c.force_, dtype=np.float64) for c in self.components])
        if dist.sharding_active():
            out = kernels.predict_multi_sharded(east, north, force_east, force_north, self.components[0].mindist, forces)
        else:
            out = np.empty((forces.shape[0], east.size), dtype=np.float64)
            kernels.predict_multi_b200(east, north, force_east, force_north, self.components[0].mindist, forces, out)
        dtype = east.dtype if np.issubdtype(east.dtype, np.floating) else np.float64
        return tuple(out[k].astype(dtype, copy=False).reshape(shape) for k in range(forces.shape[0]))


class Chain(BaseGridder):
    """
    Apply ``steps`` = [(name, gridder), ...] in sequence: each step is fitted to the
    residuals of the previous ones (through ``filter``); ``predict`` sums the steps
    that can predict.
    """

    def __init__(self, steps):
        super().__init__()
        self.steps = steps

    @property
    def named_steps(self):
        "Access the steps by name"
        return dict(self.steps)

    def fit(self, coordinates, data, weights=None):
        self.region_ = get_region(coordinates[:2])
        args = coordinates, data, weights
        for _, step in self.steps:
            args = step.filter(*args)
        return self

    def predict(self, coordinates):
        check_is_fitted(self, ["region_"])
        result = None
        for _, step in self.steps:
            if hasattr(step, "predict"):
                predicted = check_data(step.predict(coordinates))
                if result is None:
                    result = [0 for _ in range(len(predicted))]
                for i, pred in enumerate(predicted):
                    result[i] += pred
        if len(result) == 1:
            return result[0]
        return tuple(result)
