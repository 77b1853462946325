"""
``block_split``, ``BlockReduce`` and ``BlockMean``: drop-ins for
verde/coordinates.py:848-946 and verde/blockreduce.py:36-506 whose labelling and
aggregation run on the GPU (the step in front of
the spline path -- BASELINE config 3 takes its 50k force positions from the
block centres of 400k points).

What the reference does with a KD-tree (``tree.query`` of every point against
the block centres) and a pandas ``groupby().aggregate(reduction)`` becomes, in
``libverde_b200.so`` (``vb_block.cu``): a label kernel, a stable radix sort of
(block, point) pairs and one warp per non-empty block that reduces its points in
their original order.

The reference accepts ANY Python callable as *reduction* (pandas calls it once
per block on the host).  A device path can only offer a closed set; it covers
the reductions Verde's documentation and tests use -- ``np.mean`` / ``"mean"``,
``np.average`` (weighted when weights are given), ``np.median``, ``np.sum``,
``np.min``, ``np.max``, ``"count"``, ``"var"`` / ``"std"`` (pandas' ddof=1) and the callables ``np.var`` /
``np.std`` (ddof=0: pandas >= 3 applies a numpy callable as it is, which is what the pinned reference
environment does) -- and raises ``NotImplementedError`` for anything else (no CPU fallback).

NaN values: every device reduction propagates NaN (a block containing one gives NaN).  pandas routes some
numpy callables to its own NaN-skipping methods (``np.mean``, ``np.sum``, ``np.min``, ``np.max`` under
pandas 3) and not others (``np.median``); that version-dependent behaviour is not reproduced.
"""
import ctypes

import numpy as np
from sklearn.base import BaseEstimator

from . import _cabi
from .base import check_coordinates, check_fit_input
from .coords import get_region, grid_coordinates

MEAN, SUM, MIN, MAX, COUNT, MEDIAN, VAR, WMEAN, WVAR, INV_SUM_W, VAR0, STD, STD0 = range(13)

_BY_NAME = {"mean": MEAN, "average": MEAN, "sum": SUM, "min": MIN, "amin": MIN, "max": MAX, "amax": MAX,
            "count": COUNT, "size": COUNT, "median": MEDIAN, "var": VAR, "std": STD}
# numpy callables whose pandas-name twin means something else (ddof = 1): applied as they are (ddof = 0)
_NUMPY_CALLABLE = {"var": VAR0, "std": STD0}


def reduction_code(reduction, weighted=False):
    "Map a reduction callable / name onto the device reductions (see the module docstring)."
    name = reduction if isinstance(reduction, str) else getattr(reduction, "__name__", None)
    module = getattr(reduction, "__module__", "") or ""
    if not isinstance(reduction, str) and not module.startswith("numpy"):
        name = None
    if weighted:
        # the reference calls reduction(values, weights=w): of numpy's reductions only average takes weights
        if name == "average":
            return WMEAN
        raise TypeError(
            "weights were given, so the reduction must accept a 'weights' argument: only numpy.average "
            "does among the device reductions (got {!r})".format(reduction)
        )
    if not isinstance(reduction, str) and name in _NUMPY_CALLABLE:
        return _NUMPY_CALLABLE[name]
    if name in _BY_NAME:
        return _BY_NAME[name]
    raise NotImplementedError(
        "reduction {!r} has no device implementation (available: numpy mean, average, median, sum, min, max, var, "
        "std or the pandas names 'mean', 'median', 'sum', 'min', 'max', 'count', 'var', 'std'); arbitrary Python "
        "callables would need the host groupby this path replaces".format(reduction)
    )


class _Split:
    "One ``vb200_block_split`` whose sorted (block, point) order is re-used for several columns."

    def __init__(self, coordinates, spacing=None, adjust="spacing", region=None, shape=None, want_labels=False):
        lib = _cabi.require_device()
        coordinates = check_coordinates(coordinates)[:2]
        if region is None:
            region = get_region(coordinates)
        centre_east, centre_north = grid_coordinates(region, spacing=spacing, shape=shape, adjust=adjust,
                                                     pixel_register=True, meshgrid=False)
        self.block_coords = tuple(np.ravel(c) for c in np.meshgrid(centre_east, centre_north))
        east, north = (_cabi.f64(c) for c in coordinates)
        ce, cn = _cabi.f64(centre_east), _cabi.f64(centre_north)
        self.n = east.size
        self.labels = np.empty(self.n, dtype=np.int64) if want_labels else None
        nblocks = ctypes.c_int64(0)
        _cabi.check(lib.vb200_block_split(_cabi.ptr(east), _cabi.ptr(north), self.n, _cabi.ptr(ce), ce.size,
                                          _cabi.ptr(cn), cn.size, _cabi.ptr(self.labels), ctypes.byref(nblocks)),
                    "block_split")
        self.nblocks = int(nblocks.value)
        self._lib = lib

    def ids(self):
        ids = np.empty(self.nblocks, dtype=np.int64)
        _cabi.check(self._lib.vb200_block_ids(_cabi.ptr(ids)), "block_ids")
        return ids

    def aggregate(self, values, op, weights=None):
        v = None if values is None else _cabi.f64(values)
        w = None if weights is None else _cabi.f64(weights)
        for a in (v, w):
            if a is not None and a.size != self.n:
                raise ValueError("expected {} values per column, got {}".format(self.n, a.size))
        out = np.empty(self.nblocks, dtype=np.float64)
        _cabi.check(self._lib.vb200_block_aggregate(_cabi.ptr(v), _cabi.ptr(w), self.n, int(op), _cabi.ptr(out)),
                    "block_aggregate")
        return out


_STREAM_OPS = (MEAN, SUM, COUNT, WMEAN)
_STREAM_MIN_BLOCKS = 4096  # BlockReduce(order_free=True) takes the ordered path on coarser block grids


def reduce_order_free(coordinates, columns, op, weights=None, spacing=None, adjust="spacing", region=None, shape=None):
    """
    One-pass, order-free block reduction (``vb200_block_reduce_stream``): every point is read once, labelled exactly
    as ``block_split`` labels it and added into per-block accumulators with atomics -- no sort, 24 B per point and
    column instead of ~68.  *columns*: sequence of value arrays reduced together; *op*: MEAN, SUM, COUNT or WMEAN
    (with *weights*).  Returns ``(block_coords, ids, values)``: the centres of ALL blocks as ``block_split`` gives
    them, the labels of the non-empty blocks in ascending order, and one reduced array per column.  The sums differ
    from the ordered path by the rounding order of the additions only (not reproducible to the last bit).
    """
    if op not in _STREAM_OPS:
        raise NotImplementedError("the order-free path offers mean, sum, count and weighted mean")
    lib = _cabi.require_device()
    coordinates = check_coordinates(coordinates)[:2]
    if region is None:
        region = get_region(coordinates)
    centre_east, centre_north = grid_coordinates(region, spacing=spacing, shape=shape, adjust=adjust,
                                                 pixel_register=True, meshgrid=False)
    block_coords = tuple(np.ravel(c) for c in np.meshgrid(centre_east, centre_north))
    east, north = (_cabi.f64(c) for c in coordinates)
    ce, cn = _cabi.f64(centre_east), _cabi.f64(centre_north)
    cols = [_cabi.f64(c) for c in columns]
    for c in cols:
        if c.size != east.size:
            raise ValueError("expected {} values per column, got {}".format(east.size, c.size))
    matrix = np.ascontiguousarray(np.stack(cols)) if cols else None
    w = None if weights is None else _cabi.f64(weights)
    capacity = int(min(east.size, ce.size * cn.size))
    ids = np.empty(capacity, dtype=np.int64)
    out = np.empty((len(cols), capacity), dtype=np.float64)
    nblocks = ctypes.c_int64(0)
    _cabi.check(lib.vb200_block_reduce_stream(_cabi.ptr(east), _cabi.ptr(north), east.size, _cabi.ptr(ce), ce.size,
                                              _cabi.ptr(cn), cn.size, _cabi.ptr(matrix), len(cols), _cabi.ptr(w), int(op),
                                              capacity, _cabi.ptr(ids), _cabi.ptr(out), ctypes.byref(nblocks)),
                "block_reduce_stream")
    nb = int(nblocks.value)
    return block_coords, ids[:nb], [out[k, :nb] for k in range(len(cols))]


def block_split(coordinates, spacing=None, adjust="spacing", region=None, shape=None):
    """
    Split a region into blocks and label points by the block they fall in
    (verde/coordinates.py:848-946).  Returns ``(block_coordinates, labels)``:
    the centres of ALL blocks (raveled, easting fastest) and one int64 label per
    point.  Points exactly half-way between two centres go to the lower block.
    """
    split = _Split(coordinates, spacing=spacing, adjust=adjust, region=region, shape=shape, want_labels=True)
    return split.block_coords, split.labels


def variance_to_weights(variance, tol=1e-15, dtype="float64"):
    "verde/utils.py:105-178: weights = min(variance) / variance in (0, 1], 1 where the variance is ~0 or NaN."
    variance = variance if isinstance(variance, tuple) else (variance,)
    weights = []
    for var in variance:
        var = np.nan_to_num(np.atleast_1d(var), copy=True)
        w = np.ones_like(var, dtype=dtype)
        nonzero = var > tol
        if np.any(nonzero):
            nonzero_var = var[nonzero]
            w[nonzero] = nonzero_var.min() / nonzero_var
        weights.append(w)
    return weights[0] if len(weights) == 1 else tuple(weights)


class BlockReduce(BaseEstimator):
    """
    Apply a reduction to the data in blocks (verde/blockreduce.py:36-243).  Same
    parameters, same return values (``filter`` gives ``(block_coordinates,
    block_data)``, blocks without points are left out, blocks ordered by label);
    see the module docstring for the reductions the device offers.
    """

    def __init__(self, reduction, spacing=None, region=None, adjust="spacing", center_coordinates=False,
                 shape=None, drop_coords=True, order_free=False):
        self.reduction = reduction
        self.shape = shape
        self.spacing = spacing
        self.region = region
        self.adjust = adjust
        self.center_coordinates = center_coordinates
        self.drop_coords = drop_coords
        # not a parameter of the reference: True takes the one-pass atomic path for mean / sum / count / weighted
        # average (same blocks and labels; sums in arbitrary instead of pandas' order); other reductions ignore it
        self.order_free = order_free

    def _split(self, coordinates):
        return _Split(coordinates, spacing=self.spacing, shape=self.shape, adjust=self.adjust, region=self.region)

    def filter(self, coordinates, data, weights=None):  # noqa: A003
        "Blocked aggregation of *data* (and of the coordinates, unless ``center_coordinates``)."
        coordinates, data, weights = check_fit_input(coordinates, data, weights, unpack=False)
        if self.order_free:
            streamed = self._filter_order_free(coordinates, data, weights)
            if streamed is not None:
                return streamed
        split = self._split(coordinates)
        if any(w is None for w in weights):
            op = reduction_code(self.reduction)
            blocked_data = tuple(split.aggregate(comp, op) for comp in data)
        else:
            op = reduction_code(self.reduction, weighted=True)
            blocked_data = tuple(split.aggregate(comp, op, w) for comp, w in zip(data, weights))
        blocked_coords = self._block_coordinates(coordinates, split)
        if len(blocked_data) == 1:
            return blocked_coords, blocked_data[0]
        return blocked_coords, blocked_data

    def _filter_order_free(self, coordinates, data, weights):
        "``filter`` through the one-pass path, or None when the reduction (or per-component weights) needs the ordered one."
        weighted = not any(w is None for w in weights)
        op = reduction_code(self.reduction, weighted=weighted)
        coord_op = reduction_code(self.reduction)
        if op not in _STREAM_OPS or coord_op not in _STREAM_OPS:
            return None
        if weighted and any(w is not weights[0] and not np.array_equal(w, weights[0]) for w in weights[1:]):
            return None  # one weight vector per launch
        kwargs = dict(spacing=self.spacing, shape=self.shape, adjust=self.adjust, region=self.region)
        region = self.region if self.region is not None else get_region(coordinates[:2])
        centres = grid_coordinates(region, spacing=self.spacing, shape=self.shape, adjust=self.adjust,
                                   pixel_register=True, meshgrid=False)
        if centres[0].size * centres[1].size < _STREAM_MIN_BLOCKS:
            return None  # few blocks = every atomic hits the same few addresses (20M points on 8 x 8 blocks: 5.1 ms
            #              against 0.9 ms for the ordered path, profiles/r02p_blocks_mask_throughput.jsonl)
        coords = coordinates[:2] if self.drop_coords else coordinates
        extra = [] if self.center_coordinates else list(coords)
        if self.center_coordinates and not self.drop_coords:
            extra = list(coords[2:])
        if weighted:
            # the reference reduces the coordinates WITHOUT the weights (verde/blockreduce.py:214-227): second pass
            block_coords, ids, blocked = reduce_order_free(coordinates, data, op, weights[0], **kwargs)
            reduced = reduce_order_free(coordinates, extra, coord_op, None, **kwargs)[2] if extra else []
        else:
            block_coords, ids, both = reduce_order_free(coordinates, list(data) + extra, op, None, **kwargs)
            blocked, reduced = both[:len(data)], both[len(data):]
        if self.center_coordinates:
            blocked_coords = tuple(bc[ids] for bc in block_coords[:2]) + tuple(reduced)
        else:
            blocked_coords = tuple(reduced)
        blocked_data = tuple(blocked)
        return blocked_coords, (blocked_data[0] if len(blocked_data) == 1 else blocked_data)

    def _block_coordinates(self, coordinates, split):
        "Coordinates of the non-empty blocks: block centres or the (unweighted) reduction of the points."
        if self.drop_coords:
            coordinates = coordinates[:2]
        op = reduction_code(self.reduction)
        reduced = [None] * len(coordinates)
        first = 0
        if self.center_coordinates:
            unique = split.ids()
            for i, block_coord in enumerate(split.block_coords[:2]):
                reduced[i] = block_coord[unique]
            first = 2
        for i in range(first, len(coordinates)):
            reduced[i] = split.aggregate(coordinates[i], op)
        return tuple(reduced)


class BlockMean(BlockReduce):
    """
    (Weighted) block mean with output weights (verde/blockreduce.py:246-506):
    ``filter`` returns ``(block_coordinates, block_mean, block_weights)``; the
    weights come from the block variance (no input weights), from uncertainty
    propagation (``uncertainty=True``: 1/sum(w)) or from the weighted variance.
    """

    def __init__(self, spacing=None, region=None, adjust="spacing", center_coordinates=False, uncertainty=False,
                 shape=None, drop_coords=True):
        super().__init__(reduction=np.average, shape=shape, spacing=spacing, region=region, adjust=adjust,
                         center_coordinates=center_coordinates, drop_coords=drop_coords)
        self.uncertainty = uncertainty

    def filter(self, coordinates, data, weights=None):  # noqa: A003
        coordinates, data, weights = check_fit_input(coordinates, data, weights, unpack=False)
        if any(w is None for w in weights) and self.uncertainty:
            raise ValueError(
                "Weights are required for uncertainty propagation."
                "Either provide weights (as 1/uncertainty**2) or use "
                "'uncertainty=False' to produce variance weights instead."
            )
        split = self._split(coordinates)
        ncomps = len(data)
        if any(w is None for w in weights):
            mean = [split.aggregate(comp, MEAN) for comp in data]
            variance = [split.aggregate(comp, VAR) for comp in data]
        elif self.uncertainty:
            mean = [split.aggregate(comp, WMEAN, w) for comp, w in zip(data, weights)]
            variance = [split.aggregate(None, INV_SUM_W, w) for w in weights]
        else:
            mean = [split.aggregate(comp, WMEAN, w) for comp, w in zip(data, weights)]
            variance = [split.aggregate(comp, WVAR, w) for comp, w in zip(data, weights)]
        blocked_data = tuple(np.ravel(comp) for comp in mean)
        blocked_weights = tuple(variance_to_weights(np.ravel(var)) for var in variance)
        blocked_coords = self._block_coordinates(coordinates, split)
        if ncomps == 1:
            return blocked_coords, blocked_data[0], blocked_weights[0]
        return blocked_coords, blocked_data, blocked_weights
