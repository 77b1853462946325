"""
verde_b200 -- B200-native implementation of Verde's Green's-function spline
gridding hot path (``Spline``, ``SplineCV``, ``VectorSpline2D``).

The estimator API follows fatiando/verde; the numerics run in hand-written
sm_100a CUDA kernels behind the C ABI of ``include/verde_b200.h``
(``libverde_b200.so``, loaded with ctypes).  There is no CPU fallback.
"""
from . import dist  # noqa: F401
from ._cabi import EngineError  # noqa: F401
from .base import BaseGridder, GridResult, check_fit_input, n_1d_arrays, score_estimator  # noqa: F401
from .blockreduce import BlockMean, BlockReduce, block_split, variance_to_weights  # noqa: F401
from .checkerboard import CheckerBoard  # noqa: F401
from .compose import Chain, Vector  # noqa: F401
from .coords import (  # noqa: F401
    get_region,
    grid_coordinates,
    line_coordinates,
    profile_coordinates,
    scatter_points,
)
from .elastic_spline import VectorSpline2D  # noqa: F401
from .kernels import least_squares  # noqa: F401
from .mask import distance_mask  # noqa: F401
from .scalar_spline import Spline, SplineCV, cross_val_score, fit_score, select  # noqa: F401

__version__ = "0.1.0"
