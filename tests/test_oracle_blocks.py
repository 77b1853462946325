"""
Pins the CPU restatement of block_split / BlockReduce / BlockMean / distance_mask
(oracle/block_oracle.py) to outputs of the unmodified reference (tests/golden/blocks_small.npz,
oracle/make_golden_blocks.py), and checks the host-side mapping of reductions.
"""
import numpy as np
import pytest

from conftest import load_golden


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
def test_block_split_matches_reference(g, bo, tag, kw):
    blocks, labels = bo.block_split((g["east"], g["north"]), **kw)
    np.testing.assert_array_equal(blocks[0], g["split_%s_be" % tag])
    np.testing.assert_array_equal(blocks[1], g["split_%s_bn" % tag])
    np.testing.assert_array_equal(labels, g["split_%s_labels" % tag])


@pytest.mark.parametrize("tag,red", [("mean", np.mean), ("median", np.median), ("sum", np.sum), ("min", np.min), ("max", np.max)])
def test_block_reduce_matches_reference(g, bo, tag, red):
    _, labels = bo.block_split((g["east"], g["north"]), spacing=500)
    for col, key in ((g["data"], "d"), (g["east"], "e"), (g["north"], "n")):
        _, got = bo.group_reduce(labels, col, red)
        np.testing.assert_allclose(got, g["br_%s_%s" % (tag, key)], rtol=1e-14, atol=0)


def test_weighted_average_and_block_mean_match_reference(g, bo):
    _, labels = bo.block_split((g["east"], g["north"]), spacing=400)
    _, d0 = bo.group_reduce(labels, g["data"], np.average, g["weights"])
    _, d1 = bo.group_reduce(labels, g["data2"], np.average, g["weights2"])
    _, h = bo.group_reduce(labels, g["height"], np.average)
    np.testing.assert_allclose(d0, g["br_wavg_d0"], rtol=1e-14)
    np.testing.assert_allclose(d1, g["br_wavg_d1"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(h, g["br_wavg_h"], rtol=1e-14)
    _, labels = bo.block_split((g["east"], g["north"]), spacing=500)
    mean, w = bo.block_mean(labels, g["data"])
    np.testing.assert_allclose(mean, g["bm_plain_d"], rtol=1e-14)
    np.testing.assert_allclose(w, g["bm_plain_w"], rtol=1e-12)
    mean, w = bo.block_mean(labels, g["data"], g["weights"], uncertainty=True)
    np.testing.assert_allclose(mean, g["bm_unc_d"], rtol=1e-14)
    np.testing.assert_allclose(w, g["bm_unc_w"], rtol=1e-13)
    for k, (d, wt) in enumerate(((g["data"], g["weights"]), (g["data2"], g["weights2"]))):
        mean, w = bo.block_mean(labels, d, wt)
        np.testing.assert_allclose(mean, g["bm_wvar_d%d" % k], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(w, g["bm_wvar_w%d" % k], rtol=1e-12)


def test_distance_mask_matches_reference(g, bo):
    data = (g["mask_data_e"], g["mask_data_n"])
    grid = (g["mask_grid_e"], g["mask_grid_n"])
    for tag, maxdist in (("50", 50.0), ("200", 200.0), ("1000", 1000.0), ("0", 0.0)):
        got = bo.distance_mask(data, maxdist, grid)
        np.testing.assert_array_equal(got, g["mask_" + tag])
        assert got.shape == grid[0].shape
    np.testing.assert_array_equal(bo.distance_mask((2.5, -7.5), 2, (g["mask_doc_e"], g["mask_doc_n"])), g["mask_doc"])
    # nodes exactly at maxdist (3-4-5 triangles) are inside, as in the reference
    tie = bo.distance_mask(([0.0], [0.0]), 5.0, (g["mask_tie_e"], g["mask_tie_n"]))
    np.testing.assert_array_equal(tie, g["mask_tie"])
    assert tie[6 + 3, 6 + 4] and tie[6 - 5, 6] and not tie[6 + 4, 6 + 4]


def test_reduction_mapping_and_errors():
    from verde_b200 import blockreduce as br

    assert br.reduction_code(np.mean) == br.MEAN and br.reduction_code("mean") == br.MEAN
    assert br.reduction_code(np.average) == br.MEAN and br.reduction_code(np.average, weighted=True) == br.WMEAN
    assert br.reduction_code(np.median) == br.MEDIAN and br.reduction_code(np.sum) == br.SUM
    assert br.reduction_code(np.min) == br.MIN and br.reduction_code(np.amax) == br.MAX
    assert br.reduction_code("var") == br.VAR and br.reduction_code("count") == br.COUNT
    with pytest.raises(NotImplementedError):
        br.reduction_code(lambda x: x.mean())
    # pandas >= 3 applies a numpy callable as it is: np.var / np.std are ddof = 0, the names "var" / "std" ddof = 1
    assert br.reduction_code(np.var) == br.VAR0 and br.reduction_code(np.std) == br.STD0
    assert br.reduction_code("std") == br.STD
    with pytest.raises(TypeError):
        br.reduction_code(np.mean, weighted=True)
    w = br.variance_to_weights(np.array([0, 2, 0.2, 1e-16, np.nan]))
    np.testing.assert_allclose(w, [1, 0.1, 1, 1, 1])


def test_block_paths_fail_loudly_without_cuda():
    import verde_b200 as vb

    if vb._cabi.load().vb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(vb.EngineError):
        vb.BlockReduce(np.mean, spacing=1).filter(([0.0, 1.0], [0.0, 1.0]), [1.0, 2.0])
    with pytest.raises(vb.EngineError):
        vb.distance_mask(([0.0], [0.0]), 1.0, coordinates=([0.0], [0.0]))
