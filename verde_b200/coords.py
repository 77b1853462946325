# Host glue derived from fatiando/verde (BSD 3-Clause): it restates the reference's estimator contract, error and
# warning texts so that this path is a drop-in.
#   Copyright (c) 2017 The Verde Developers.  Distributed under the terms of the BSD 3-Clause License.
#   SPDX-License-Identifier: BSD-3-Clause
#   This code is part of the Fatiando a Terra project (https://www.fatiando.org); see LICENSE-verde.txt.
"""
Coordinate generators that define the input / grid layout of the spline path.

Host-side numpy only (no arithmetic worth accelerating).  Behaviour follows
verde/coordinates.py: ``check_region`` (:19-50), ``get_region`` (:53-81),
``scatter_points`` (:120-190), ``line_coordinates`` (:192-290),
``grid_coordinates`` (:292-591), ``spacing_to_size`` (:593-640),
``profile_coordinates`` (:698-767).
"""
import numpy as np
from sklearn.utils import check_random_state


def check_region(region):
    "Raise ValueError unless region = (W, E, S, N) with W <= E and S <= N."
    if len(region) != 4:
        raise ValueError("Invalid region '{}'. Only 4 values allowed.".format(region))
    west, east, south, north = region
    if west > east:
        raise ValueError(
            "Invalid region '{}' (W, E, S, N). Must have W =< E. ".format(region)
            + "If working with geographic coordinates, don't forget to match geographic"
            + " region with coordinates using 'verde.longitude_continuity'."
        )
    if south > north:
        raise ValueError("Invalid region '{}' (W, E, S, N). Must have S =< N.".format(region))


def get_region(coordinates):
    "Bounding box (W, E, S, N) of the first two coordinate arrays."
    easting, northing = coordinates[:2]
    return (np.min(easting), np.max(easting), np.min(northing), np.max(northing))


def _with_extra(coordinates, extra_coords):
    if extra_coords is not None:
        for value in np.atleast_1d(extra_coords):
            coordinates.append(np.ones_like(coordinates[0]) * value)
    return tuple(coordinates)


def scatter_points(region, size, random_state=None, extra_coords=None):
    "Uniform random points: easting drawn first, then northing, from one RandomState."
    check_region(region)
    rng = check_random_state(random_state)
    bounds = np.array(region).reshape((len(region) // 2, 2))
    coordinates = [rng.uniform(lower, upper, size) for lower, upper in bounds]
    return _with_extra(coordinates, extra_coords)


def spacing_to_size(start, stop, spacing, adjust):
    "Number of nodes (and possibly adjusted stop) for a given spacing."
    if adjust not in ("spacing", "region"):
        raise ValueError(
            "Invalid value for *adjust* '{}'. Should be 'spacing' or 'region'".format(adjust)
        )
    size = int(round((stop - start) / spacing)) + 1
    if size == 1:
        size += 1
    if adjust == "region":
        stop = start + (size - 1) * spacing
    return size, stop


def line_coordinates(start, stop, size=None, spacing=None, adjust="spacing", pixel_register=False):
    "Evenly spaced values (np.linspace semantics), grid-line or pixel registered."
    if size is not None and spacing is not None:
        raise ValueError("Both size and spacing provided. Only one is allowed.")
    if size is None and spacing is None:
        raise ValueError("Either a size or a spacing must be provided.")
    if spacing is not None:
        size, stop = spacing_to_size(start, stop, spacing, adjust)
    elif pixel_register:
        size = size + 1
    values = np.linspace(start, stop, size)
    if pixel_register:
        values = values[:-1] + (values[1] - values[0]) / 2
    return values


def grid_coordinates(region, shape=None, spacing=None, adjust="spacing", pixel_register=False,
                     extra_coords=None, meshgrid=True):
    """
    Regular grid over *region*; shape = (n_north, n_east).  With ``meshgrid=True``
    returns C-ordered 2-D arrays (k = i_north * n_east + i_east when raveled).
    """
    check_region(region)
    if shape is not None and spacing is not None:
        raise ValueError("Both grid shape and spacing provided. Only one is allowed.")
    if shape is None and spacing is None:
        raise ValueError("Either a grid shape or a spacing must be provided.")
    if shape is None:
        shape = (None, None)
        spacing = np.atleast_1d(spacing)
        if len(spacing) == 1:
            spacing = (spacing[0], spacing[0])
        elif len(spacing) > 2:
            raise ValueError(f"Only two values allowed for grid spacing: {spacing}")
    else:
        spacing = (None, None)
    east = line_coordinates(region[0], region[1], size=shape[1], spacing=spacing[1], adjust=adjust,
                            pixel_register=pixel_register)
    north = line_coordinates(region[2], region[3], size=shape[0], spacing=spacing[0], adjust=adjust,
                             pixel_register=pixel_register)
    if not meshgrid:
        if extra_coords is not None:
            raise ValueError(
                "Using 'meshgrid=False' and 'extra_coords' is 'verde.grid_coordinates' is not allowed."
            )
        return (east, north)
    return _with_extra(list(np.meshgrid(east, north)), extra_coords)


def profile_coordinates(point1, point2, size, extra_coords=None):
    "Points along the straight line point1 -> point2 and their distances from point1."
    if size <= 0:
        raise ValueError("Invalid profile size '{}'. Must be > 0.".format(size))
    diffs = [b - a for a, b in zip(point1, point2)]
    separation = np.hypot(*diffs)
    distances = np.linspace(0, separation, size)
    angle = np.arctan2(*reversed(diffs))
    coordinates = [point1[0] + distances * np.cos(angle), point1[1] + distances * np.sin(angle)]
    return _with_extra(coordinates, extra_coords), distances
