"""
Golden fixtures for the block reductions and the distance mask, from the UNMODIFIED reference
(verde.block_split / BlockReduce / BlockMean / distance_mask on pandas + scipy cKDTree).

    python oracle/make_golden_blocks.py        (build container only: needs /root/reference)

TEST INFRASTRUCTURE ONLY.  Writes tests/golden/blocks_small.npz.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refimport import import_reference  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    warnings.simplefilter("ignore")
    vd = import_reference()
    import pandas
    import scipy

    out = dict(numpy_version=np.__version__, pandas_version=pandas.__version__, scipy_version=scipy.__version__)
    region = (0, 5000, -5000, 0)
    tab = vd.synthetic.CheckerBoard(region=region).scatter(size=3000, random_state=3)
    east, north, data = tab.easting.values, tab.northing.values, tab.scalars.values
    rng = np.random.RandomState(7)
    data2 = rng.normal(size=east.size)
    height = rng.uniform(0, 100, east.size)
    weights = rng.uniform(0.1, 10, east.size)
    weights2 = rng.uniform(0.5, 2, east.size)
    out.update(east=east, north=north, data=data, data2=data2, height=height, weights=weights, weights2=weights2)

    # block_split: spacing (adjusted), shape, explicit region larger than the data
    for tag, kw in (("sp500", dict(spacing=500)), ("sh7x9", dict(shape=(7, 9))),
                    ("sp330_region", dict(spacing=330, region=(-500, 5600, -5200, 300), adjust="region")),
                    ("sp300x700", dict(spacing=(300, 700)))):
        blocks, labels = vd.block_split((east, north), **kw)
        out["split_%s_be" % tag], out["split_%s_bn" % tag], out["split_%s_labels" % tag] = blocks[0], blocks[1], labels

    # BlockReduce: mean / median / sum / min / max, centred or reduced coordinates
    for tag, red in (("mean", np.mean), ("median", np.median), ("sum", np.sum), ("min", np.min), ("max", np.max)):
        c, d = vd.BlockReduce(red, spacing=500).filter((east, north), data)
        out["br_%s_e" % tag], out["br_%s_n" % tag], out["br_%s_d" % tag] = c[0], c[1], d
    c, d = vd.BlockReduce(np.median, shape=(7, 9), center_coordinates=True).filter((east, north), data)
    out["br_median_center_e"], out["br_median_center_n"], out["br_median_center_d"] = c[0], c[1], d
    # weighted average, two components, extra coordinate kept
    c, d = vd.BlockReduce(np.average, spacing=400, drop_coords=False).filter((east, north, height), (data, data2),
                                                                             (weights, weights2))
    out["br_wavg_e"], out["br_wavg_n"], out["br_wavg_h"] = c
    out["br_wavg_d0"], out["br_wavg_d1"] = d

    # BlockMean: the three weighting modes
    c, d, w = vd.BlockMean(spacing=500).filter((east, north), data)
    out["bm_plain_e"], out["bm_plain_n"], out["bm_plain_d"], out["bm_plain_w"] = c[0], c[1], d, w
    c, d, w = vd.BlockMean(spacing=500, uncertainty=True).filter((east, north), data, weights)
    out["bm_unc_d"], out["bm_unc_w"] = d, w
    c, d, w = vd.BlockMean(spacing=500, center_coordinates=True).filter((east, north), (data, data2), (weights, weights2))
    out["bm_wvar_e"], out["bm_wvar_n"] = c[0], c[1]
    out["bm_wvar_d0"], out["bm_wvar_d1"], out["bm_wvar_w0"], out["bm_wvar_w1"] = d[0], d[1], w[0], w[1]

    # BASELINE config 3: force positions = block centres of 400k points on a 224 x 224 block grid
    big = vd.synthetic.CheckerBoard(region=region).scatter(size=400_000, random_state=0)
    c, d = vd.BlockReduce(np.mean, shape=(224, 224), center_coordinates=True).filter(
        (big.easting.values, big.northing.values), big.scalars.values)
    out["cfg3_n_blocks"] = np.int64(c[0].size)
    out["cfg3_head_e"], out["cfg3_head_n"], out["cfg3_head_d"] = c[0][:200], c[1][:200], d[:200]
    out["cfg3_sums"] = np.array([c[0].sum(), c[1].sum(), d.sum(), np.abs(d).sum()])

    # distance_mask
    data_e, data_n = vd.scatter_points((500, 4200, -4600, -300), size=2000, random_state=11)
    grid_e, grid_n = vd.grid_coordinates(region, shape=(100, 120))
    out.update(mask_data_e=data_e, mask_data_n=data_n, mask_grid_e=grid_e, mask_grid_n=grid_n)
    for tag, maxdist in (("50", 50.0), ("200", 200.0), ("1000", 1000.0), ("0", 0.0)):
        out["mask_" + tag] = vd.distance_mask((data_e, data_n), maxdist=maxdist, coordinates=(grid_e, grid_n))
    # the docstring example (verde/mask.py:75-88), a single data point; distances of exactly 2 do not occur
    coords = vd.grid_coordinates((0, 5, -10, -4), spacing=1)
    out["mask_doc_e"], out["mask_doc_n"] = coords
    out["mask_doc"] = vd.distance_mask((2.5, -7.5), maxdist=2, coordinates=coords)
    # nodes EXACTLY at maxdist: 3-4-5 triangles around an integer data point
    ee, nn = np.meshgrid(np.arange(-6.0, 7.0), np.arange(-6.0, 7.0))
    out["mask_tie_e"], out["mask_tie_n"] = ee, nn
    out["mask_tie"] = vd.distance_mask((np.array([0.0]), np.array([0.0])), maxdist=5.0, coordinates=(ee, nn))
    np.savez_compressed(os.path.join(GOLDEN, "blocks_small.npz"), **out)
    print("wrote blocks_small.npz:", {k: np.shape(v) for k, v in out.items() if not k.endswith("version")})


if __name__ == "__main__":
    main()
