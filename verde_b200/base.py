# Host glue derived from fatiando/verde (BSD 3-Clause): it restates the reference's estimator contract, error and
# warning texts so that this path is a drop-in.
#   Copyright (c) 2017 The Verde Developers.  Distributed under the terms of the BSD 3-Clause License.
#   SPDX-License-Identifier: BSD-3-Clause
#   This code is part of the Fatiando a Terra project (https://www.fatiando.org); see LICENSE-verde.txt.
"""
Thin restatement of the estimator contract of ``verde.base.BaseGridder``.

This is the drop-in boundary: subclasses implement
``fit(coordinates, data, weights=None) -> self`` and ``predict(coordinates)``;
``filter / score / grid / scatter / profile`` are inherited and only ever call
those two.  Contract, names, defaults, warnings and errors follow
verde/base/base_classes.py:135-829 and verde/base/utils.py:15-295; nothing here
is a kernel target, it is host glue so that a user of ``verde.Spline`` finds the
same methods on the B200 gridders.

``grid()`` returns an ``xarray.Dataset`` when xarray is importable (as the
reference does, verde/utils.py:373-522) and otherwise a :class:`GridResult`
with the same ``coords`` / ``data_vars`` / ``attrs`` content, because this
image ships no xarray.
"""
import warnings

import numpy as np
import pandas as pd
from sklearn.base import BaseEstimator
from sklearn.metrics import check_scoring

from .coords import grid_coordinates, profile_coordinates, scatter_points


# --------------------------------------------------------------------------
# input validation (verde/base/utils.py:93-295)
# --------------------------------------------------------------------------
def check_data(data):
    "Always a tuple of data components."
    return data if isinstance(data, tuple) else (data,)


def check_data_names(data, data_names):
    "Validate *data_names* against the number of data components."
    if isinstance(data_names, str):
        data_names = (data_names,)
    if data_names is None:
        raise ValueError("Invalid data_names equal to None.")
    if len(data) != len(data_names):
        raise ValueError(
            "Data has {} components but only {} names provided: {}".format(
                len(data), len(data_names), str(data_names)
            )
        )
    return data_names


def check_coordinates(coordinates):
    "All coordinate arrays must share one shape."
    shapes = [np.shape(c) for c in coordinates]
    if any(shape != shapes[0] for shape in shapes):
        raise ValueError(
            "Coordinate arrays must have the same shape. Coordinate shapes: {}".format(shapes)
        )
    return coordinates


def check_extra_coords_names(coordinates, extra_coords_names):
    "Validate names for coordinates beyond (easting, northing)."
    if isinstance(extra_coords_names, str):
        extra_coords_names = (extra_coords_names,)
    if extra_coords_names is None:
        raise ValueError(
            "Invalid extra_coords_names equal to None. "
            + "When passing one or more extra coordinate, "
            + "extra_coords_names cannot be None."
        )
    if len(coordinates[2:]) != len(extra_coords_names):
        raise ValueError(
            "Invalid extra_coords_names '{}'. ".format(extra_coords_names)
            + "Number of extra coordinates names must match the number of "
            + "additional coordinates ('{}').".format(len(coordinates[2:]))
        )
    return extra_coords_names


def check_fit_input(coordinates, data, weights, unpack=True):
    "Shape checks shared by every ``fit``; weights are raveled (verde/base/utils.py:205-263)."
    data = check_data(data)
    weights = check_data(weights)
    coordinates = check_coordinates(coordinates)
    ref_shape = np.shape(coordinates[0])
    if any(np.shape(d) != ref_shape for d in data):
        raise ValueError(
            "Data arrays must have the same shape {} as coordinates. Data shapes: {}.".format(
                ref_shape, [np.shape(d) for d in data]
            )
        )
    if any(w is not None for w in weights):
        if len(weights) != len(data):
            raise ValueError(
                "Number of data '{}' and weights '{}' must be equal.".format(len(data), len(weights))
            )
        if any(np.size(w) != np.size(d) for w in weights for d in data):
            raise ValueError("Weights must have the same size as the data array.")
        weights = tuple(np.ravel(w) for w in weights)
    else:
        weights = tuple([None] * len(data))
    if unpack:
        if len(weights) == 1:
            weights = weights[0]
        if len(data) == 1:
            data = data[0]
    return coordinates, data, weights


def n_1d_arrays(arrays, n):
    "First *n* arrays as raveled (C-order) 1-D numpy arrays."
    return tuple(np.ravel(np.atleast_1d(a)) for a in arrays[:n])


class DummyEstimator(BaseEstimator):
    "Carries already-computed predictions into scikit-learn's scorer API."

    def __init__(self, predicted):
        self._predicted = predicted

    def predict(self, *args, **kwargs):  # noqa: U100
        return self._predicted

    def fit(self, *args, **kwargs):  # noqa: U100
        return self


def score_estimator(scoring, estimator, coordinates, data, weights=None):
    "Mean over data components of a scikit-learn scorer (verde/base/utils.py:15-66)."
    coordinates, data, weights = check_fit_input(coordinates, data, weights, unpack=False)
    predicted = check_data(estimator.predict(coordinates))
    scorer = check_scoring(DummyEstimator, scoring=scoring)
    return np.mean(
        [
            scorer(DummyEstimator(np.ravel(pred)), coordinates, np.ravel(data[i]), sample_weight=weights[i])
            for i, pred in enumerate(predicted)
        ]
    )


# --------------------------------------------------------------------------
# grid packaging
# --------------------------------------------------------------------------
class GridResult:
    """
    Minimal stand-in for ``xarray.Dataset`` (used only when xarray is missing).

    ``coords`` maps dimension/extra-coordinate names to arrays, ``data_vars`` maps
    data names to ``(dims, 2-D array)``; ``result[name]`` returns the 2-D array.
    """

    def __init__(self, data_vars, coords, dims):
        self.data_vars = dict(data_vars)
        self.coords = dict(coords)
        self.dims = tuple(dims)
        self.attrs = {}

    def __getitem__(self, name):
        if name in self.data_vars:
            return self.data_vars[name][1]
        return self.coords[name]

    def __iter__(self):
        return iter(self.data_vars)

    def where(self, mask):
        "Like ``xarray.Dataset.where``: a copy with every data variable set to NaN where *mask* is False."
        mask = np.asarray(mask, dtype=bool)
        data_vars = {name: (dims, np.where(mask, np.asarray(values, dtype=np.float64), np.nan))
                     for name, (dims, values) in self.data_vars.items()}
        out = GridResult(data_vars, self.coords, self.dims)
        out.attrs = dict(self.attrs)
        return out

    def __repr__(self):
        shape = next(iter(self.data_vars.values()))[1].shape if self.data_vars else ()
        return "<GridResult dims={} shape={} data_vars={}>".format(self.dims, shape, list(self.data_vars))


def get_ndim_horizontal_coords(easting, northing):
    ndim = np.ndim(easting)
    if ndim != np.ndim(northing):
        raise ValueError(
            "Horizontal coordinates dimensions mismatch. "
            + f"The easting coordinate array has {np.ndim(easting)} dimensions "
            + f"while the northing has {np.ndim(northing)}."
        )
    return ndim


def check_meshgrid(coordinates):
    easting, northing = coordinates[:2]
    msg = (
        "Invalid coordinate array. The arrays for the horizontal "
        + "coordinates of a regular grid must be meshgrids."
    )
    if not np.allclose(easting[0, :], easting):
        raise ValueError(msg)
    if not np.allclose(northing[:, 0][:, None], northing):
        raise ValueError(msg)


def meshgrid_from_1d(coordinates):
    ndim = get_ndim_horizontal_coords(*coordinates[:2])
    if ndim != 1:
        raise ValueError("Horizontal coordinates must be 1d-arrays. " + f"{ndim}d-arrays provided.")
    easting, northing = np.meshgrid(coordinates[0], coordinates[1])
    return check_coordinates((easting, northing, *coordinates[2:]))


def make_grid(coordinates, data, data_names, dims=("northing", "easting"), extra_coords_names=None):
    "Package gridded predictions (xarray.Dataset if available, else GridResult)."
    if get_ndim_horizontal_coords(*coordinates[:2]) == 2:
        check_coordinates(coordinates)
        check_meshgrid(coordinates)
        coordinates = (coordinates[0][0, :], coordinates[1][:, 0], *coordinates[2:])
    coords = {dims[1]: coordinates[0], dims[0]: coordinates[1]}
    if coordinates[2:]:
        extra_coords_names = check_extra_coords_names(coordinates, extra_coords_names)
        for name, extra in zip(extra_coords_names, coordinates[2:]):
            coords[name] = (dims, extra)
    data_vars = None
    if data is not None:
        data = check_data(data)
        data_names = check_data_names(data, data_names)
        data_vars = {name: (dims, value) for name, value in zip(data_names, data)}
    try:
        import xarray as xr
    except ImportError:
        return GridResult(data_vars or {}, coords, dims)
    return xr.Dataset(data_vars, coords)


def project_coordinates(coordinates, projection, **kwargs):
    projected = projection(*coordinates[:2], **kwargs)
    if len(coordinates) > 2:
        projected += tuple(coordinates[2:])
    return projected


def get_instance_region(instance, region):
    if region is None:
        if not hasattr(instance, "region_"):
            raise ValueError("No default region found. Argument must be supplied.")
        region = instance.region_
    return region


# --------------------------------------------------------------------------
# the estimator base class
# --------------------------------------------------------------------------
class BaseGridder(BaseEstimator):
    """
    scikit-learn style base for gridders (see the module docstring).

    ``__init__`` of subclasses must only assign parameters (so that
    ``sklearn.base.clone`` / ``get_params`` / ``set_params`` work, which
    ``cross_val_score`` relies on); fitted state ends in ``_``.
    """

    dims = ("northing", "easting")
    extra_coords_name = "extra_coord"
    data_names_defaults = [
        ("scalars",),
        ("east_component", "north_component"),
        ("east_component", "north_component", "vertical_component"),
    ]

    def predict(self, coordinates):  # noqa: U100
        raise NotImplementedError()

    def fit(self, coordinates, data, weights=None):  # noqa: U100
        raise NotImplementedError()

    def filter(self, coordinates, data, weights=None):  # noqa: A003
        "Fit, then return (coordinates, data - prediction, weights)."
        self.fit(coordinates, data, weights)
        data = check_data(data)
        pred = check_data(self.predict(coordinates))
        residuals = tuple(d - p.reshape(np.shape(d)) for d, p in zip(data, pred))
        if len(residuals) == 1:
            residuals = residuals[0]
        return coordinates, residuals, weights

    def score(self, coordinates, data, weights=None):
        "R² of the predictions (mean over components); warns like the reference."
        warnings.warn(
            "The default scoring will change from R² to negative root mean "
            "squared error (RMSE) in Verde 2.0.0. "
            "This may change model selection results slightly.",
            FutureWarning,
            stacklevel=2,
        )
        return score_estimator("r2", self, coordinates, data, weights=weights)

    def grid(self, region=None, shape=None, spacing=None, dims=None, data_names=None,
             projection=None, coordinates=None, **kwargs):
        "Predict on a regular grid (region defaults to the fitted ``region_``)."
        if coordinates is not None and (spacing is not None or shape is not None):
            raise ValueError(
                "Both coordinates and spacing or shape were provided. "
                + "Please pass only coordinates or either the spacing or the shape."
            )
        if coordinates is not None and region is not None:
            raise ValueError(
                "Both coordinates and region were provided. "
                + "Please pass region only if spacing or shape is specified."
            )
        if coordinates is not None:
            if get_ndim_horizontal_coords(*coordinates[:2]) == 1:
                coordinates = meshgrid_from_1d(coordinates)
            else:
                check_meshgrid(coordinates)
        else:
            region = get_instance_region(self, region)
            coordinates = grid_coordinates(region, shape=shape, spacing=spacing, **kwargs)
        if projection is None:
            data = check_data(self.predict(coordinates))
        else:
            data = check_data(self.predict(project_coordinates(coordinates, projection)))
        dims = self._get_dims(dims)
        data_names = self._get_data_names(data, data_names)
        extra_coords_names = self._get_extra_coords_names(coordinates)
        dataset = make_grid(coordinates, data, data_names, dims=dims, extra_coords_names=extra_coords_names)
        metadata = "Generated by {}".format(repr(self))
        dataset.attrs["metadata"] = metadata
        if not isinstance(dataset, GridResult):
            for array in dataset:
                dataset[array].attrs["metadata"] = metadata
        return dataset

    def _table(self, coordinates, data, dims, data_names, leading=()):
        data_names = self._get_data_names(data, data_names)
        columns = [(dims[0], coordinates[1]), (dims[1], coordinates[0])]
        columns.extend(leading)
        columns.extend(zip(self._get_extra_coords_names(coordinates), coordinates[2:]))
        columns.extend(zip(data_names, data))
        return pd.DataFrame(dict(columns), columns=[c[0] for c in columns])

    def scatter(self, region=None, size=300, random_state=0, dims=None, data_names=None,
                projection=None, **kwargs):
        "Predict on a random scatter (deprecated in the reference; same warning)."
        warnings.warn(
            "The 'scatter' method is deprecated and will be removed in Verde "
            "2.0.0. Use 'verde.scatter_points' and the 'predict' method "
            "instead.",
            FutureWarning,
            stacklevel=2,
        )
        dims = self._get_dims(dims)
        region = get_instance_region(self, region)
        coordinates = scatter_points(region, size, random_state=random_state, **kwargs)
        if projection is None:
            data = check_data(self.predict(coordinates))
        else:
            data = check_data(self.predict(project_coordinates(coordinates, projection)))
        return self._table(coordinates, data, dims, data_names)

    def profile(self, point1, point2, size, dims=None, data_names=None, projection=None, **kwargs):
        "Predict along a straight profile; distances are Cartesian."
        dims = self._get_dims(dims)
        if projection is not None:
            point1 = project_coordinates(point1, projection)
            point2 = project_coordinates(point2, projection)
        coordinates, distances = profile_coordinates(point1, point2, size, **kwargs)
        data = check_data(self.predict(coordinates))
        if projection is not None:
            coordinates = project_coordinates(coordinates, projection, inverse=True)
        return self._table(coordinates, data, dims, data_names, leading=[("distance", distances)])

    def _get_dims(self, dims):
        return self.dims if dims is None else dims

    def _get_extra_coords_names(self, coordinates):
        names = []
        for i in range(len(coordinates[2:])):
            name = self.extra_coords_name
            if i > 0:
                name += "_{}".format(i)
            names.append(name)
        return names

    def _get_data_names(self, data, data_names):
        if data_names is None:
            if len(data) > len(self.data_names_defaults):
                raise ValueError(
                    "Default data names only available for up to 3 components. "
                    + "Must provide custom names through the 'data_names' argument."
                )
            return self.data_names_defaults[len(data) - 1]
        return check_data_names(data, data_names)
