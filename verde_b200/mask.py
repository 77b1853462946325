"""
``distance_mask``: drop-in for verde/mask.py:17-113 (SURVEY.md section 8f, rank 4:
the step that follows ``grid()`` in the gallery examples), on the GPU.

The reference builds a KD-tree of the data points, asks for the nearest-neighbour
distance of every grid node and keeps ``distance <= maxdist``.  Here the data
points are binned into a uniform cell grid of pitch >= maxdist on the device and
every node looks for a point within ``maxdist`` in its 3 x 3 cell neighbourhood
(``vb200_distance_mask``); the Boolean result is identical, including nodes at
exactly ``maxdist`` (same unfused ``sqrt(dx*dx + dy*dy)``).
"""
import numpy as np

from . import _cabi
from .base import GridResult, check_coordinates, n_1d_arrays


def _get_grid_coordinates(coordinates, grid):
    "verde/mask.py:200-234: coordinates either given or taken from the grid's first two dimensions."
    if coordinates is None and grid is None:
        raise ValueError("Either coordinates or grid must be given.")
    if coordinates is None:
        if isinstance(grid, GridResult):
            dims = grid.dims
        else:
            dims = [grid[var].dims for var in grid.data_vars][0]
        coordinates = np.meshgrid(np.asarray(grid.coords[dims[1]]), np.asarray(grid.coords[dims[0]]))
    check_coordinates(coordinates)
    shape = np.shape(coordinates[0])
    return coordinates, shape


def distance_mask(data_coordinates, maxdist, coordinates=None, grid=None, projection=None):
    """
    True where a point is at most *maxdist* from the closest data point.

    With *coordinates*: a Boolean array of their shape.  With *grid* (an
    ``xarray.Dataset`` or this package's ``GridResult``): the grid with values
    outside the mask set to NaN (``grid.where(mask)``).
    """
    lib = _cabi.require_device()
    coordinates, shape = _get_grid_coordinates(coordinates, grid)
    if projection is not None:
        data_coordinates = projection(*n_1d_arrays(data_coordinates, 2))
        coordinates = projection(*n_1d_arrays(coordinates, 2))
    data_east, data_north = (_cabi.f64(c) for c in n_1d_arrays(data_coordinates, 2))
    east, north = (_cabi.f64(c) for c in n_1d_arrays(coordinates, 2))
    if data_east.size != data_north.size:
        raise ValueError("data coordinate arrays must have the same size")
    mask = np.zeros(east.size, dtype=np.uint8)
    _cabi.check(lib.vb200_distance_mask(_cabi.ptr(data_east), _cabi.ptr(data_north), data_east.size, float(maxdist),
                                        _cabi.ptr(east), _cabi.ptr(north), east.size, _cabi.ptr(mask)),
                "distance_mask")
    mask = mask.astype(bool).reshape(shape)
    if grid is not None:
        return grid.where(mask)
    return mask
