"""
Host-side logic of the estimator layer that needs no GPU: constructor / clone
contract, warnings and error types of the reference,
coordinate generators, input validation.
"""
import warnings

import numpy as np
import pytest
from sklearn.base import clone
from sklearn.exceptions import NotFittedError

import _refimport
import verde_b200 as vb


def test_constructor_signatures_and_clone():
    with pytest.warns(FutureWarning, match="mindist parameter of verde.Spline"):
        sp = vb.Spline(mindist=1e-5, damping=1e-6, force_coords=([1.0], [2.0]))
    assert sp.get_params() == {"mindist": 1e-5, "damping": 1e-6, "force_coords": ([1.0], [2.0]), "engine": "auto"}
    assert vb.Spline().mindist == 0 and vb.Spline().damping is None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", FutureWarning)
        twin = clone(sp)
    assert twin.get_params() == sp.get_params() and twin is not sp
    vs = vb.VectorSpline2D()
    assert vs.get_params() == {"poisson": 0.5, "mindist": 10e3, "damping": None, "force_coords": None, "engine": "auto"}
    cv = vb.SplineCV()
    assert cv.mindists == [0] and cv.dampings == (1e-10, 1e-5, 1e-1) and cv.cv is None and cv.scoring is None
    # BlockReduce: the reference's parameters plus the opt-in one-pass flavour (off by default -> a drop-in)
    br = vb.BlockReduce(np.mean, spacing=2.0)
    assert br.get_params() == {"reduction": np.mean, "spacing": 2.0, "region": None, "adjust": "spacing",
                               "center_coordinates": False, "shape": None, "drop_coords": True, "order_free": False}
    assert clone(vb.BlockReduce(np.sum, shape=(3, 4), order_free=True)).order_free is True
    assert vb.BlockMean(spacing=2.0).get_params()["uncertainty"] is False and vb.BlockMean(spacing=2.0).order_free is False
    from verde_b200 import blockreduce

    assert blockreduce.reduction_code(np.mean) in blockreduce._STREAM_OPS
    assert blockreduce.reduction_code(np.median) not in blockreduce._STREAM_OPS
    assert blockreduce.reduction_code(np.average, weighted=True) == blockreduce.WMEAN
    with pytest.raises(NotImplementedError):
        blockreduce.reduction_code(lambda v: v.max() - v.min())


def test_deprecation_warnings_match_reference():
    with pytest.warns(FutureWarning, match="'engine' parameter of 'verde.Spline'"):
        vb.Spline(engine="numba")
    with pytest.warns(FutureWarning, match="'engine' parameter of 'verde.VectorSpline2D'"):
        vb.VectorSpline2D(engine="numpy")
    with pytest.warns(FutureWarning, match="mindists parameter of verde.SplineCV"):
        vb.SplineCV(mindists=[1.0])
    with pytest.warns(FutureWarning, match="'client' parameter of 'verde.SplineCV'"):
        vb.SplineCV(client=object())


def test_fit_input_validation_errors():
    e = np.linspace(0, 1, 8)
    with pytest.raises(ValueError, match="Coordinate arrays must have the same shape"):
        vb.Spline(damping=1.0).fit((e, e[:5]), e)
    with pytest.raises(ValueError, match="Data arrays must have the same shape"):
        vb.Spline(damping=1.0).fit((e, e), e[:5])
    with pytest.raises(ValueError, match="Weights must have the same size"):
        vb.Spline(damping=1.0).fit((e, e), e, weights=e[:3])
    with pytest.raises(ValueError, match="Need two data components"):
        vb.VectorSpline2D(damping=1.0).fit((e, e), (e, e, e))
    with pytest.raises(ValueError, match="Invalid engine"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", FutureWarning)
            vb.Spline(damping=1.0, engine="cuda?").fit((e, e), e)


def test_weights_without_damping_warns_like_reference():
    e = np.linspace(0, 1, 8)
    with pytest.warns(UserWarning, match="Weights might have no effect if no regularization is used"):
        try:
            vb.Spline().fit((e, e), e, weights=np.ones(8))  # damping=None: the truncated-SVD solve (GPU only)
        except vb.EngineError:
            pass  # no CUDA device here: the warning must still have been issued before the engine is reached


def test_predict_before_fit_raises_not_fitted():
    e = np.linspace(0, 1, 4)
    for est in (vb.Spline(damping=1.0), vb.VectorSpline2D(damping=1.0), vb.SplineCV()):
        with pytest.raises(NotFittedError):
            est.predict((e, e))


def test_grid_coordinates_layout():
    east, north = vb.grid_coordinates((0, 10, -5, 5), shape=(3, 6))
    assert east.shape == (3, 6) and north.shape == (3, 6)
    np.testing.assert_array_equal(east[0], np.linspace(0, 10, 6))
    np.testing.assert_array_equal(north[:, 0], np.linspace(-5, 5, 3))
    # raveled index k = i_north * n_east + i_east (C order), the layout the grid predict shards
    assert east.ravel()[1 * 6 + 2] == east[1, 2] and north.ravel()[1 * 6 + 2] == north[1, 2]
    e1, n1 = vb.grid_coordinates((0, 10, -5, 5), spacing=2.5)
    assert e1.shape == (5, 5)
    e2, _ = vb.grid_coordinates((0, 10, -5, 5), shape=(2, 5), pixel_register=True)
    np.testing.assert_allclose(e2[0], [1, 3, 5, 7, 9])
    with pytest.raises(ValueError, match="Either a grid shape or a spacing"):
        vb.grid_coordinates((0, 1, 0, 1))
    with pytest.raises(ValueError, match="Must have W =< E"):
        vb.grid_coordinates((1, 0, 0, 1), shape=(2, 2))


def test_scatter_points_and_checkerboard_are_the_reference_generators(golden_cfg1):
    tab = vb.CheckerBoard().scatter(size=5000, random_state=0)
    assert list(tab.columns) == ["northing", "easting", "scalars"]
    np.testing.assert_array_equal(tab.easting.values[:16], golden_cfg1["east_head"])
    np.testing.assert_array_equal(tab.northing.values[:16], golden_cfg1["north_head"])
    np.testing.assert_array_equal(tab.scalars.values[:16], golden_cfg1["data_head"])
    assert vb.get_region((tab.easting, tab.northing)) == (tab.easting.min(), tab.easting.max(),
                                                          tab.northing.min(), tab.northing.max())


def test_checkerboard_grid_and_profile_need_no_gpu():
    cb = vb.CheckerBoard(region=(0, 100, 0, 50))
    grid = cb.grid(shape=(5, 7))
    assert np.asarray(grid["scalars"]).shape == (5, 7)
    assert "CheckerBoard" in grid.attrs["metadata"]
    prof = cb.profile((0, 0), (100, 50), 11)
    assert list(prof.columns) == ["northing", "easting", "distance", "scalars"]
    np.testing.assert_allclose(prof.distance.values[-1], np.hypot(100, 50))
    with pytest.raises(ValueError, match="Both coordinates and region"):
        cb.grid(region=(0, 1, 0, 1), coordinates=(np.arange(3.0), np.arange(3.0)))


def test_select_and_score_estimator():
    a = (np.arange(5), np.arange(-3, 2))
    out = vb.select(a, [1, 3])
    np.testing.assert_array_equal(out[0], [1, 3])
    np.testing.assert_array_equal(out[1], [-2, 0])
    assert vb.select((None, None), [0]) == (None, None)

    class Const:
        def predict(self, coordinates):
            return np.full(np.shape(coordinates[0]), 2.0)

    e = np.arange(4.0)
    score = vb.score_estimator("neg_root_mean_squared_error", Const(), (e, e), np.array([2.0, 2.0, 4.0, 0.0]))
    np.testing.assert_allclose(score, -np.sqrt(2.0))


@pytest.mark.skipif(not _refimport.reference_available(), reason="reference checkout only exists in the build container")
def test_coordinate_helpers_against_live_reference():
    vd = _refimport.import_reference()
    for kwargs in (dict(shape=(7, 9)), dict(spacing=0.3), dict(spacing=(0.5, 0.2), adjust="region"),
                   dict(shape=(4, 4), pixel_register=True), dict(shape=(3, 5), extra_coords=[10, 20])):
        ours = vb.grid_coordinates((-2, 3, 1, 4), **kwargs)
        theirs = vd.grid_coordinates((-2, 3, 1, 4), **kwargs)
        assert len(ours) == len(theirs)
        for a, b in zip(ours, theirs):
            np.testing.assert_array_equal(a, b)
    for a, b in zip(vb.scatter_points((0, 5, -1, 1), 50, random_state=3, extra_coords=7),
                    vd.scatter_points((0, 5, -1, 1), 50, random_state=3, extra_coords=7)):
        np.testing.assert_array_equal(a, b)
    (c1, d1), (c2, d2) = vb.profile_coordinates((1, 2), (5, 9), 13), vd.profile_coordinates((1, 2), (5, 9), 13)
    np.testing.assert_array_equal(d1, d2)
    np.testing.assert_array_equal(c1[0], c2[0])
