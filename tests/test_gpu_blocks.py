"""
Device block reductions (block_split / BlockReduce / BlockMean) and distance_mask against the reference
goldens (tests/golden/blocks_small.npz) and the CPU oracle (oracle/block_oracle.py).

Bars: labels, block ids, counts and masks are integer / Boolean work -> bit exact.  Sums are floating point:
1e-13 relative to the block's magnitude (the device adds in the same order as pandas, compensated).
"""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-13


@pytest.fixture(scope="module")
def vb():
    import verde_b200

    verde_b200._cabi.require_device()
    return verde_b200


@pytest.fixture(scope="module")
def g():
    return load_golden("blocks_small.npz")


@pytest.fixture(scope="module")
def bo():
    import block_oracle

    return block_oracle


SPLITS = [("sp500", dict(spacing=500)), ("sh7x9", dict(shape=(7, 9))),
          ("sp330_region", dict(spacing=330, region=(-500, 5600, -5200, 300), adjust="region")),
          ("sp300x700", dict(spacing=(300, 700)))]


@pytest.mark.parametrize("tag,kw", SPLITS)
def test_block_split_matches_reference(vb, g, tag, kw):
    blocks, labels = vb.block_split((g["east"], g["north"]), **kw)
    np.testing.assert_array_equal(blocks[0], g["split_%s_be" % tag])
    np.testing.assert_array_equal(blocks[1], g["split_%s_bn" % tag])
    np.testing.assert_array_equal(labels, g["split_%s_labels" % tag])
    assert labels.dtype == np.int64


@pytest.mark.parametrize("tag,red", [("mean", np.mean), ("median", np.median), ("sum", np.sum), ("min", np.min), ("max", np.max)])
def test_block_reduce_matches_reference(vb, g, tag, red):
    coords, data = vb.BlockReduce(red, spacing=500).filter((g["east"], g["north"]), g["data"])
    np.testing.assert_allclose(data, g["br_%s_d" % tag], rtol=RTOL, atol=RTOL * np.abs(g["data"]).max())
    np.testing.assert_allclose(coords[0], g["br_%s_e" % tag], rtol=RTOL)
    np.testing.assert_allclose(coords[1], g["br_%s_n" % tag], rtol=RTOL)
    if tag in ("median", "min", "max"):  # selections, not sums: bit exact
        np.testing.assert_array_equal(data, g["br_%s_d" % tag])


def test_block_reduce_center_coordinates_and_components(vb, g):
    coords, data = vb.BlockReduce(np.median, shape=(7, 9), center_coordinates=True).filter((g["east"], g["north"]), g["data"])
    np.testing.assert_array_equal(coords[0], g["br_median_center_e"])
    np.testing.assert_array_equal(coords[1], g["br_median_center_n"])
    np.testing.assert_array_equal(data, g["br_median_center_d"])
    coords, data = vb.BlockReduce(np.average, spacing=400, drop_coords=False).filter(
        (g["east"], g["north"], g["height"]), (g["data"], g["data2"]), (g["weights"], g["weights2"]))
    assert len(coords) == 3 and isinstance(data, tuple) and len(data) == 2
    for got, key in zip(coords + data, ("br_wavg_e", "br_wavg_n", "br_wavg_h", "br_wavg_d0", "br_wavg_d1")):
        np.testing.assert_allclose(got, g[key], rtol=1e-12, atol=1e-14)


def test_block_mean_three_weightings(vb, g):
    coords, mean, w = vb.BlockMean(spacing=500).filter((g["east"], g["north"]), g["data"])
    np.testing.assert_allclose(mean, g["bm_plain_d"], rtol=RTOL, atol=1e-10)
    np.testing.assert_allclose(w, g["bm_plain_w"], rtol=1e-11)
    np.testing.assert_allclose(coords[0], g["bm_plain_e"], rtol=RTOL)
    np.testing.assert_allclose(coords[1], g["bm_plain_n"], rtol=RTOL)
    _, mean, w = vb.BlockMean(spacing=500, uncertainty=True).filter((g["east"], g["north"]), g["data"], g["weights"])
    np.testing.assert_allclose(mean, g["bm_unc_d"], rtol=RTOL, atol=1e-10)
    np.testing.assert_allclose(w, g["bm_unc_w"], rtol=1e-12)
    coords, mean, w = vb.BlockMean(spacing=500, center_coordinates=True).filter(
        (g["east"], g["north"]), (g["data"], g["data2"]), (g["weights"], g["weights2"]))
    np.testing.assert_array_equal(coords[0], g["bm_wvar_e"])
    np.testing.assert_array_equal(coords[1], g["bm_wvar_n"])
    for k in range(2):
        np.testing.assert_allclose(mean[k], g["bm_wvar_d%d" % k], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(w[k], g["bm_wvar_w%d" % k], rtol=1e-10)
    with pytest.raises(ValueError):
        vb.BlockMean(spacing=500, uncertainty=True).filter((g["east"], g["north"]), g["data"])


def test_config3_force_positions(vb, g):
    "BASELINE config 3: block centres of 400 000 CheckerBoard points on a 224 x 224 block grid."
    tab = vb.CheckerBoard(region=(0, 5000, -5000, 0)).scatter(size=400_000, random_state=0)
    coords, data = vb.BlockReduce(np.mean, shape=(224, 224), center_coordinates=True).filter(
        (tab.easting.values, tab.northing.values), tab.scalars.values)
    assert coords[0].size == int(g["cfg3_n_blocks"])
    np.testing.assert_array_equal(coords[0][:200], g["cfg3_head_e"])
    np.testing.assert_array_equal(coords[1][:200], g["cfg3_head_n"])
    np.testing.assert_allclose(data[:200], g["cfg3_head_d"], rtol=RTOL, atol=1e-10)
    sums = np.array([coords[0].sum(), coords[1].sum(), data.sum(), np.abs(data).sum()])
    np.testing.assert_allclose(sums, g["cfg3_sums"], rtol=1e-12)


@pytest.mark.parametrize("n,shape", [(1, (1, 1)), (5, (2, 3)), (4099, (1, 1)), (70_001, (3, 2)), (20_000, (300, 400)),
                                     (100_000, (1000, 1000))])
def test_block_reduce_vs_oracle_shapes(vb, bo, n, shape):
    """Single point, one block holding everything (segments longer than a sort tile), ragged tails, label
    ranges that need 1, 2 and 3 radix passes, mostly empty block grids."""
    rng = np.random.RandomState(n)
    e, nn = rng.uniform(-3, 7, n), rng.uniform(10, 11, n)
    d = rng.normal(size=n) * 10.0 ** rng.randint(-3, 4, n)
    w = rng.uniform(0.1, 3, n)
    region = (-3, 7, 10, 11)
    blocks, labels = vb.block_split((e, nn), shape=shape, region=region)
    _, ref_labels = bo.block_split((e, nn), shape=shape, region=region)
    np.testing.assert_array_equal(labels, ref_labels)
    ids, ref_mean = bo.group_reduce(ref_labels, d, np.mean)
    for red, ref_fn in ((np.mean, np.mean), (np.median, np.median), (np.sum, np.sum), (np.max, np.max)):
        coords, got = vb.BlockReduce(red, shape=shape, region=region, center_coordinates=True).filter((e, nn), d)
        _, ref = bo.group_reduce(ref_labels, d, ref_fn)
        assert got.shape == ref.shape
        scale = np.abs(d).max() * (max(1, n // ids.size) if red is np.sum else 1)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-13 * scale)
        np.testing.assert_array_equal(coords[0], blocks[0][ids])
    _, got = vb.BlockReduce(np.average, shape=shape, region=region).filter((e, nn), d, w)
    _, ref = bo.group_reduce(ref_labels, d, np.average, w)
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-13 * np.abs(d).max())
    _, got = vb.BlockReduce("count", shape=shape, region=region).filter((e, nn), d)
    np.testing.assert_array_equal(got, np.bincount(ref_labels)[ids])
    # numpy callables are applied as they are (ddof = 0, pandas >= 3); the pandas names mean ddof = 1
    small = d / np.abs(d).max()
    for red, ref_fn in ((np.var, np.var), (np.std, np.std), ("var", bo.sample_variance),
                        ("std", lambda v: np.sqrt(bo.sample_variance(v)))):
        _, got = vb.BlockReduce(red, shape=shape, region=region).filter((e, nn), small)
        _, ref = bo.group_reduce(ref_labels, small, ref_fn)
        np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-15)


@pytest.mark.parametrize("n,shape", [(1, (1, 1)), (4099, (1, 1)), (70_001, (3, 2)), (20_000, (300, 400)), (400_000, (224, 224))])
def test_order_free_reduction_matches_the_ordered_path(vb, bo, g, n, shape):
    """BlockReduce(order_free=True): one pass, atomics into per-block accumulators (vb200_block_reduce_stream).  Blocks,
    labels and counts are the ordered path's bit for bit; sums may differ by the rounding order of the additions."""
    rng = np.random.RandomState(n + 7)
    e, nn, h = rng.uniform(-3, 7, n), rng.uniform(10, 11, n), rng.uniform(0, 100, n)
    d, d2 = rng.normal(size=n) * 10.0 ** rng.randint(-3, 4, n), rng.uniform(-1, 1, n)
    w = rng.uniform(0.1, 3, n)
    region = (-3, 7, 10, 11)
    kw = dict(shape=shape, region=region)
    vb.blockreduce._STREAM_MIN_BLOCKS = 0  # coarse grids too (BlockReduce would route them to the ordered path)
    for red, centre in ((np.mean, True), (np.mean, False), (np.sum, False), ("count", True)):
        want_c, want = vb.BlockReduce(red, center_coordinates=centre, **kw).filter((e, nn), (d, d2))
        got_c, got = vb.BlockReduce(red, center_coordinates=centre, order_free=True, **kw).filter((e, nn), (d, d2))
        assert len(got) == 2 and len(got_c) == 2
        per_block = max(1, n // want[0].size)
        for a, b in zip(got + got_c, want + want_c):
            assert a.shape == b.shape
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-13 * np.abs(d).max() * (per_block if red is np.sum else 1))
        if centre:  # block centres of the non-empty blocks: selections, bit exact
            np.testing.assert_array_equal(got_c[0], want_c[0])
            np.testing.assert_array_equal(got_c[1], want_c[1])
        if red == "count":
            np.testing.assert_array_equal(got[0], want[0])
    # weighted average, a third coordinate kept (reduced WITHOUT the weights, as the reference does)
    want_c, want = vb.BlockReduce(np.average, drop_coords=False, **kw).filter((e, nn, h), d, w)
    got_c, got = vb.BlockReduce(np.average, drop_coords=False, order_free=True, **kw).filter((e, nn, h), d, w)
    assert len(got_c) == 3
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13 * np.abs(d).max())
    for a, b in zip(got_c, want_c):
        np.testing.assert_allclose(a, b, rtol=1e-12)
    # against the numpy oracle too
    _, ref_labels = bo.block_split((e, nn), **kw)
    ids, ref = bo.group_reduce(ref_labels, d, np.mean)
    _, ids_got, vals = vb.blockreduce.reduce_order_free((e, nn), [d], vb.blockreduce.MEAN, **kw)
    np.testing.assert_array_equal(ids_got, ids)
    np.testing.assert_allclose(vals[0], ref, rtol=1e-12, atol=1e-13 * np.abs(d).max())
    # reductions the one-pass path does not offer take the ordered path, unchanged; so do coarse block grids
    vb.blockreduce._STREAM_MIN_BLOCKS = 4096
    a = vb.BlockReduce(np.median, order_free=True, **kw).filter((e, nn), d)[1]
    b = vb.BlockReduce(np.median, **kw).filter((e, nn), d)[1]
    np.testing.assert_array_equal(a, b)
    if shape[0] * shape[1] < 4096:
        a = vb.BlockReduce(np.mean, order_free=True, **kw).filter((e, nn), d)[1]
        b = vb.BlockReduce(np.mean, **kw).filter((e, nn), d)[1]
        np.testing.assert_array_equal(a, b)


def test_order_free_reduction_abi_errors(vb):
    import ctypes

    lib = vb._cabi.require_device()
    e = np.linspace(0, 1, 50)
    ce, cn = np.array([0.25, 0.75]), np.array([0.5])
    ids, out, nb = np.empty(1, dtype=np.int64), np.empty(1), ctypes.c_int64()
    args = (e.ctypes.data, e.ctypes.data, 50, ce.ctypes.data, 2, cn.ctypes.data, 1, e.ctypes.data, 1, None)
    assert lib.vb200_block_reduce_stream(*args, vb.blockreduce.MEAN, 1, ids.ctypes.data, out.ctypes.data, ctypes.byref(nb)) != 0
    assert nb.value == 2  # the capacity was too small: the count comes back, the call fails
    assert lib.vb200_block_reduce_stream(*args, vb.blockreduce.MEDIAN, 2, ids.ctypes.data, out.ctypes.data, ctypes.byref(nb)) != 0
    assert lib.vb200_block_reduce_stream(*args, vb.blockreduce.WMEAN, 2, ids.ctypes.data, out.ctypes.data, ctypes.byref(nb)) != 0


def test_block_reduce_edge_cases(vb, bo):
    # points outside the region go to the nearest (edge) block, like the KD-tree query
    e = np.array([-50.0, 0.2, 0.7, 1.4, 99.0, 0.2])
    n = np.array([0.1, -80.0, 0.6, 0.4, 0.9, 55.0])
    blocks, labels = vb.block_split((e, n), spacing=0.5, region=(0, 2, 0, 1))
    _, ref = bo.block_split((e, n), spacing=0.5, region=(0, 2, 0, 1))
    np.testing.assert_array_equal(labels, ref)
    # NaN data: mean/sum/min/max/median of a block with a NaN are NaN (numpy semantics), other blocks unaffected
    d = np.array([1.0, 2.0, np.nan, 4.0, 5.0, 6.0])
    for red in (np.mean, np.median, np.min, np.max, np.sum):
        _, got = vb.BlockReduce(red, spacing=0.5, region=(0, 2, 0, 1)).filter((e, n), d)
        _, want = bo.group_reduce(ref, d, red)
        np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
        np.testing.assert_allclose(got[~np.isnan(want)], want[~np.isnan(want)], rtol=1e-15)
    # 2-D inputs are raveled; a single-point variance is NaN -> weight 1
    _, mean, w = vb.BlockMean(spacing=0.5, region=(0, 2, 0, 1)).filter((e.reshape(2, 3), n.reshape(2, 3)), np.arange(6.0).reshape(2, 3))
    assert mean.ndim == 1 and np.all(w <= 1) and np.all(w > 0)
    # unsupported reductions are refused, not run on the host
    with pytest.raises(NotImplementedError):
        vb.BlockReduce(lambda x: x.max() - x.min(), spacing=0.5).filter((e, n), d)
    with pytest.raises(ValueError):
        vb.BlockReduce(np.mean, spacing=0.5).filter((e, n[:3]), d)
    # determinism
    rng = np.random.RandomState(0)
    e, n, d = rng.uniform(0, 1, 50_000), rng.uniform(0, 1, 50_000), rng.normal(size=50_000)
    a = vb.BlockReduce(np.mean, shape=(20, 20)).filter((e, n), d)[1]
    b = vb.BlockReduce(np.mean, shape=(20, 20)).filter((e, n), d)[1]
    np.testing.assert_array_equal(a, b)


def test_distance_mask_matches_reference(vb, g):
    data = (g["mask_data_e"], g["mask_data_n"])
    grid = (g["mask_grid_e"], g["mask_grid_n"])
    for tag, maxdist in (("50", 50.0), ("200", 200.0), ("1000", 1000.0), ("0", 0.0)):
        got = vb.distance_mask(data, maxdist, coordinates=grid)
        assert got.dtype == bool and got.shape == grid[0].shape
        np.testing.assert_array_equal(got, g["mask_" + tag])
    np.testing.assert_array_equal(vb.distance_mask((2.5, -7.5), maxdist=2, coordinates=(g["mask_doc_e"], g["mask_doc_n"])),
                                  g["mask_doc"])
    np.testing.assert_array_equal(vb.distance_mask((np.array([0.0]), np.array([0.0])), maxdist=5.0,
                                                   coordinates=(g["mask_tie_e"], g["mask_tie_n"])), g["mask_tie"])


@pytest.mark.parametrize("n,maxdist", [(1, 0.3), (50, 0.01), (5000, 0.05), (5000, 3.0), (20_000, 1e-4)])
def test_distance_mask_vs_oracle(vb, bo, n, maxdist):
    rng = np.random.RandomState(n)
    de, dn = rng.uniform(0, 1, n), rng.uniform(2, 2.5, n)
    qe, qn = rng.uniform(-0.5, 1.5, (90, 70)), rng.uniform(1.5, 3.0, (90, 70))
    qe[0, :10], qn[0, :10] = de[:10] if n >= 10 else de[0], dn[:10] if n >= 10 else dn[0]  # queries ON data points
    got = vb.distance_mask((de, dn), maxdist, coordinates=(qe, qn))
    np.testing.assert_array_equal(got, bo.distance_mask((de, dn), maxdist, (qe, qn)))
    assert got[0, :10].all()


def test_distance_mask_api(vb, g):
    with pytest.raises(ValueError):
        vb.distance_mask((g["mask_data_e"], g["mask_data_n"]), 10.0)
    # projection is applied to both point sets
    got = vb.distance_mask((g["mask_data_e"], g["mask_data_n"]), 400.0, coordinates=(g["mask_grid_e"], g["mask_grid_n"]),
                           projection=lambda e, n: (2 * e, 2 * n))
    np.testing.assert_array_equal(got, g["mask_200"])
    # grid= : masks a gridded result (GridResult here, xarray.Dataset when xarray is installed)
    tab = vb.CheckerBoard().scatter(size=300, random_state=0)
    sp = vb.Spline(damping=1e-3).fit((tab.easting.values, tab.northing.values), tab.scalars.values)
    grid = sp.grid(shape=(30, 40))
    masked = vb.distance_mask((tab.easting.values, tab.northing.values), 300.0, grid=grid)
    values = np.asarray(masked["scalars"])
    coords = vb.grid_coordinates(sp.region_, shape=(30, 40))
    inside = vb.distance_mask((tab.easting.values, tab.northing.values), 300.0, coordinates=coords)
    assert np.isnan(values[~inside]).all() and np.isfinite(values[inside]).all()
    # infinite / negative maxdist
    assert vb.distance_mask((tab.easting.values, tab.northing.values), np.inf, coordinates=coords).all()
    assert not vb.distance_mask((tab.easting.values, tab.northing.values), -1.0, coordinates=coords).any()
    # empty data: nothing is close
    assert not vb.distance_mask((np.array([]), np.array([])), 1.0, coordinates=coords).any()


def test_gallery_pipeline_blockreduce_spline_mask(vb, bo):
    """doc/gallery_src/spline.py: Chain([BlockReduce(mean), Spline(damping)]) -> grid -> distance_mask, every
    numeric step on the device; checked against the same steps done one by one with the CPU oracles."""
    import verde_oracle as vo

    tab = vb.CheckerBoard(region=(0, 5000, -5000, 0)).scatter(size=6000, random_state=2)
    coords, data = (tab.easting.values, tab.northing.values), tab.scalars.values
    spacing = 250.0
    chain = vb.Chain([("mean", vb.BlockReduce(np.mean, spacing=spacing)), ("spline", vb.Spline(damping=1e-6))])
    chain.fit(coords, data)
    grid = chain.grid(spacing=100.0)
    masked = vb.distance_mask(coords, maxdist=150.0, grid=grid)
    # the same pipeline through the oracles
    _, labels = bo.block_split(coords, spacing=spacing)
    _, be = bo.group_reduce(labels, coords[0], np.mean)
    _, bn = bo.group_reduce(labels, coords[1], np.mean)
    _, bd = bo.group_reduce(labels, data, np.mean)
    force = vo.spline_fit(be, bn, bd, damping=1e-6)
    ge, gn = vb.grid_coordinates(chain.region_, spacing=100.0)
    ref = vo.predict(ge.ravel(), gn.ravel(), be, bn, 0.0, force).reshape(ge.shape)
    inside = bo.distance_mask(coords, 150.0, (ge, gn))
    got = np.asarray(masked["scalars"])
    np.testing.assert_array_equal(np.isnan(got), ~inside)
    assert np.max(np.abs(got[inside] - ref[inside])) <= 1e-6 * np.ptp(data)
    np.testing.assert_allclose(chain.named_steps["spline"].force_coords_[0], be, rtol=1e-13)
