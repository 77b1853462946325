"""
The reference's OWN behavioural tests for this path (verde/tests/test_spline.py, test_vector.py,
test_model_selection.py), restated against verde_b200 with the same data and the same tolerances.  The variants
that use the un-damped default (``damping=None``) are kept twice: with a small positive damping, where the
reference's "exact interpolation" tolerances hold, and with ``damping=None`` itself
(``test_default_constructors_behave_like_the_pinned_reference``) against what the unmodified reference produces in
the pinned environment -- under scikit-learn 1.9 ``LinearRegression`` truncates at 1e-6 sigma_max, so the
reference's own rtol=1e-5 assertions fail there too (max miss 113 on test_spline's data).
"""
import warnings

import numpy as np
import numpy.testing as npt
import pytest
from sklearn.model_selection import ShuffleSplit

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vb():
    import verde_b200

    verde_b200._cabi.require_device()
    return verde_b200


def scalars(result):
    return np.asarray(result["scalars"])


def test_spline_cv(vb):
    "verde/tests/test_spline.py:33-76 with dampings [1e-10, 1e3] instead of [None, 1e3]"
    region = (100, 500, -800, -700)
    synth = vb.CheckerBoard(region=region)
    data = synth.scatter(size=1500, random_state=0)
    coords = (data.easting, data.northing)
    spline = vb.SplineCV(dampings=[1e-10, 1e3], cv=ShuffleSplit(n_splits=1, random_state=0)).fit(coords, data.scalars)
    assert spline.damping_ == 1e-10
    # the reference asserts rtol=1e-5 for its exact solution; with damping 1e-10 the reference itself misses the
    # data by up to 3.0 here (measured with the live reference), so the damped tests' tolerance applies
    npt.assert_allclose(spline.predict(coords), data.scalars, rtol=1e-2, atol=10)
    npt.assert_allclose(spline.score(coords, data.scalars), 1, atol=1e-5)
    npt.assert_allclose(spline.force_coords_, coords)
    assert data.scalars.size == spline.force_.size
    npt.assert_allclose(spline.region_, (data.easting.min(), data.easting.max(), data.northing.min(), data.northing.max()))
    shape, sub = (5, 5), (270, 320, -770, -720)
    npt.assert_allclose(scalars(spline.grid(region=sub, shape=shape)), scalars(synth.grid(region=sub, shape=shape)), rtol=5e-2)


def test_spline_near_exact(vb):
    "verde/tests/test_spline.py:79-99 with damping 1e-10 standing in for the exact solution"
    region = (100, 500, -800, -700)
    synth = vb.CheckerBoard(region=region)
    data = synth.scatter(size=1500, random_state=1)
    coords = (data.easting, data.northing)
    spline = vb.Spline(damping=1e-10).fit(coords, data.scalars)
    npt.assert_allclose(spline.predict(coords), data.scalars, rtol=1e-2, atol=10)  # reference at 1e-10: max miss 1.24
    npt.assert_allclose(spline.score(coords, data.scalars), 1, atol=1e-5)
    assert data.scalars.size == spline.force_.size
    npt.assert_allclose(spline.force_coords_, coords)
    shape, sub = (5, 5), (270, 320, -770, -720)
    npt.assert_allclose(scalars(spline.grid(region=sub, shape=shape)), scalars(synth.grid(region=sub, shape=shape)), rtol=5e-2)


def test_spline_cv_scoring(vb):
    "verde/tests/test_spline.py:102-117: SplineCV with one parameter set == cross_val_score of that Spline"
    synth = vb.CheckerBoard(region=(100, 500, -800, -700))
    data = synth.scatter(size=1500, random_state=1)
    coords = (data.easting, data.northing)
    for score in ["r2", "neg_root_mean_squared_error"]:
        score_spline = np.mean(vb.cross_val_score(vb.Spline(damping=1e-6), coords, data.scalars, scoring=score))
        spline_cv = vb.SplineCV(dampings=[1e-6], scoring=score).fit(coords, data.scalars)
        npt.assert_allclose(score_spline, spline_cv.scores_[0], rtol=1e-5)


def test_spline_weights(vb):
    "verde/tests/test_spline.py:120-135: a tiny weight switches an outlier off"
    data = vb.CheckerBoard().scatter(size=2000, random_state=1)
    data_outlier = data.scalars.values.copy()
    outlier, outlier_value = 500, 100e3
    data_outlier[outlier] += outlier_value
    weights = np.ones_like(data_outlier)
    weights[outlier] = 1e-10
    coords = (data.easting, data.northing)
    spline = vb.Spline(damping=1e-8).fit(coords, data_outlier, weights=weights)
    npt.assert_allclose(spline.score(coords, data.scalars), 1, atol=0.01)
    predicted = spline.predict(coords)
    npt.assert_allclose(predicted, data.scalars, rtol=1e-2, atol=10)
    npt.assert_allclose(data_outlier[outlier] - predicted[outlier], outlier_value, rtol=1e-2)


def test_spline_region(vb):
    "verde/tests/test_spline.py:138-145 (2-D meshgrid inputs; damped)"
    region = (1000, 5000, -8000, -6000)
    grid = vb.CheckerBoard(region=region).grid(shape=(10, 10))
    coords = np.meshgrid(np.asarray(grid.coords["easting"]), np.asarray(grid.coords["northing"]))
    grd = vb.Spline(damping=1e-8).fit(coords, scalars(grid))
    npt.assert_allclose(grd.region_, region)


def test_spline_damping(vb):
    "verde/tests/test_spline.py:148-166"
    region = (1000, 5000, -8000, -6000)
    synth = vb.CheckerBoard(region=region)
    data = synth.scatter(size=3000, random_state=1)
    coords = (data.easting, data.northing)
    spline = vb.Spline(damping=1e-8).fit(coords, data.scalars)
    npt.assert_allclose(spline.predict(coords), data.scalars, rtol=1e-2, atol=10)
    shape, sub = (5, 5), (2000, 4000, -7500, -6500)
    npt.assert_allclose(scalars(spline.grid(region=sub, shape=shape)), scalars(synth.grid(region=sub, shape=shape)),
                        rtol=1e-2, atol=10)


def test_spline_warns_underdetermined(vb):
    "verde/tests/test_spline.py:181-188"
    data = vb.CheckerBoard().scatter(size=50, random_state=100)
    grd = vb.Spline(damping=1e-3, force_coords=(np.arange(60), np.arange(60)))
    with warnings.catch_warnings(record=True) as warn:
        warnings.simplefilter("always")
        grd.fit((data.easting, data.northing), data.scalars)
        assert any(str(w.message).startswith("Under-determined problem") for w in warn)


def test_spline_jacobian_and_predict_implementations(vb):
    "verde/tests/test_spline.py:191-209: the reference compares its numpy and numba engines; here GPU vs CPU oracle"
    import verde_oracle as vo

    data = vb.CheckerBoard().scatter(size=1500, random_state=1)
    e, n = data.easting.values, data.northing.values
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", FutureWarning)
        jac = vb.Spline(engine="numba").jacobian((e, n), (e, n))
    npt.assert_allclose(jac, vo.jacobian(e, n, e, n, 0.0), rtol=1e-7, atol=1e-9)  # the reference's rtol
    data = vb.CheckerBoard().scatter(size=500, random_state=1)
    e, n = data.easting.values, data.northing.values
    spline = vb.Spline(damping=1e-6).fit((e, n), data.scalars)
    npt.assert_allclose(spline.predict((e, n)), vo.predict(e, n, e, n, 0.0, spline.force_), rtol=1e-7, atol=1e-7)


@pytest.fixture
def data2d(vb):
    synth = vb.CheckerBoard()
    coordinates = vb.grid_coordinates(synth.region, shape=(15, 20))
    data = tuple(synth.predict(coordinates).ravel() for i in range(2))
    return tuple(i.ravel() for i in coordinates), data


def test_vector2d(vb, data2d):
    "verde/tests/test_vector.py:31-40 (damping 1e-10 for the exact solution)"
    coords, data = data2d
    spline = vb.VectorSpline2D(damping=1e-10).fit(coords, data)
    npt.assert_allclose(spline.score(coords, data), 1, atol=1e-5)
    npt.assert_allclose(spline.predict(coords), data, rtol=1e-2, atol=1)
    assert data[0].size == spline.force_coords[0].size
    assert data[0].size * 2 == spline.force_.size
    npt.assert_allclose(spline.force_coords, vb.n_1d_arrays(coords, n=2))


def test_vector2d_weights(vb, data2d):
    "verde/tests/test_vector.py:43-56"
    coords, data = data2d
    outlier, outlier_value = 10, 100000
    data_outlier = tuple(i.copy() for i in data)
    data_outlier[0][outlier] += outlier_value
    data_outlier[1][outlier] += outlier_value
    weights = tuple(np.ones_like(data_outlier[0]) for i in range(2))
    weights[0][outlier] = 1e-10
    weights[1][outlier] = 1e-10
    spline = vb.VectorSpline2D(damping=1e-8).fit(coords, data_outlier, weights)
    npt.assert_allclose(spline.score(coords, data), 1, atol=0.01)
    npt.assert_allclose(spline.predict(coords), data, rtol=1e-2, atol=5)


def test_vector2d_fails(vb, data2d):
    "verde/tests/test_vector.py:59-66"
    coords, data = data2d
    spline = vb.VectorSpline2D(damping=1e-8)
    with pytest.raises(ValueError):
        spline.fit(coords, data[0])
    with pytest.raises(ValueError):
        spline.fit(coords, data + coords)


def test_vector2d_jacobian_and_predict_implementations(vb, data2d):
    "verde/tests/test_vector.py:69-84: GPU vs CPU oracle instead of numpy vs numba"
    import verde_oracle as vo

    coords, data = data2d
    e, n = coords
    jac = vb.VectorSpline2D().jacobian(coords, coords)
    npt.assert_allclose(jac, vo.jacobian_2d(e, n, e, n, 10e3, 0.5), rtol=1e-7)
    spline = vb.VectorSpline2D(damping=1e-8).fit(coords, data)
    pe, pn = spline.predict(coords)
    re, rn = vo.predict_2d(e, n, e, n, 10e3, 0.5, spline.force_)
    npt.assert_allclose(pe, re, atol=1e-4)
    npt.assert_allclose(pn, rn, atol=1e-4)


def test_cross_val_score_of_a_vector_of_splines(vb):
    "verde/tests/test_model_selection.py:57-66 uses Vector([Trend, Trend]); the path's gridder is Spline"
    coords = vb.grid_coordinates((0, 10, -10, -5), spacing=0.5)
    data = 10 - coords[0] + 0.5 * coords[1]
    model = vb.Vector([vb.Spline(damping=1e-10), vb.Spline(damping=1e-10)])
    scores = vb.cross_val_score(model, coords, (data, data))
    npt.assert_allclose(scores, 1, atol=1e-3)
    scores = vb.cross_val_score(vb.Spline(damping=1e-10), coords, data, cv=ShuffleSplit(n_splits=3, random_state=0))
    assert len(scores) == 3
    npt.assert_allclose(scores, 1, atol=1e-3)


def test_default_constructors_behave_like_the_pinned_reference(vb):
    """
    verde/tests/test_spline.py:33-99 and test_vector.py:31-40 with the reference's own ``damping=None``.  Expected
    values: the unmodified reference run here with scikit-learn 1.9.0 / scipy 1.18.1 (oracle/_refimport.py):
    Spline() on test_spline's data scores 0.998547768467705 (max miss 113.1); SplineCV([None, 1e3]) picks None with
    scores [9.96846484e-01, 7.80646073e-04]; VectorSpline2D() interpolates (full rank, max miss 6e-9).
    """
    region = (100, 500, -800, -700)
    synth = vb.CheckerBoard(region=region)
    data = synth.scatter(size=1500, random_state=1)
    coords = (data.easting, data.northing)
    spline = vb.Spline().fit(coords, data.scalars)
    npt.assert_allclose(spline.score(coords, data.scalars), 0.998547768467705, atol=1e-8)
    npt.assert_allclose(np.max(np.abs(spline.predict(coords) - data.scalars)), 113.11381572345215, rtol=1e-5)
    assert data.scalars.size == spline.force_.size
    data = synth.scatter(size=1500, random_state=0)
    coords = (data.easting, data.northing)
    cv = vb.SplineCV(dampings=[None, 1e3], cv=ShuffleSplit(n_splits=1, random_state=0)).fit(coords, data.scalars)
    assert cv.damping_ is None
    npt.assert_allclose(cv.scores_, [9.96846484e-01, 7.80646073e-04], rtol=1e-6)
    synth = vb.CheckerBoard()
    coordinates = vb.grid_coordinates(synth.region, shape=(15, 20))
    data2 = tuple(synth.predict(coordinates).ravel() for i in range(2))
    flat = tuple(i.ravel() for i in coordinates)
    vector = vb.VectorSpline2D().fit(flat, data2)
    npt.assert_allclose(vector.score(flat, data2), 1, atol=1e-5)
    npt.assert_allclose(vector.predict(flat), data2, rtol=1e-2, atol=1)  # the reference's tolerance
