"""
Throughput of the device block reductions and the distance mask with inputs
resident in HBM, against the HBM roofline.  One JSON object per case on stdout.
    python tools/run_blocks.py
Algorithmic bytes: block mean of one column = read easting, northing, value once
(24 B/point) ; distance mask = read the query coordinates, write one byte (17 B/node).
"""
import ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import verde_b200 as vb
from verde_b200 import _cabi

lib = _cabi.require_device()
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm_peak = float(peaks.get("hbm_gbs", 6544.0))  # measured copy bandwidth of this pool's B200s


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for n, shape in ((400_000, (224, 224)), (20_000_000, (224, 224)), (20_000_000, (2000, 2000)), (20_000_000, (8, 8))):
    rng = np.random.RandomState(0)
    e, nn, d = rng.uniform(0, 5000, n), rng.uniform(-5000, 0, n), rng.normal(size=n)
    ce = np.linspace(0, 5000, shape[1] + 1); ce = ce[:-1] + (ce[1] - ce[0]) / 2
    cn = np.linspace(-5000, 0, shape[0] + 1); cn = cn[:-1] + (cn[1] - cn[0]) / 2
    de, dn, dd, dce, dcn = dev(e), dev(nn), dev(d), dev(ce), dev(cn)
    nblocks = ctypes.c_int64()

    def split():
        _cabi.check(lib.vb200_block_split(_cabi.ptr(de), _cabi.ptr(dn), n, _cabi.ptr(dce), ce.size, _cabi.ptr(dcn), cn.size,
                                          None, ctypes.byref(nblocks)), "split")

    t_split = timed(split)
    out = torch.empty(nblocks.value, dtype=torch.float64, device="cuda")
    res = {"case": "block_reduce", "points": n, "blocks": list(shape), "non_empty": int(nblocks.value), "split_ms": t_split}
    for name, op in (("median", 5), ("mean", 0)):  # mean last: `out` keeps it for the order-free comparison below
        t = timed(lambda: _cabi.check(lib.vb200_block_aggregate(_cabi.ptr(dd), None, n, op, _cabi.ptr(out)), "agg"))
        res[name + "_ms"] = t
    total = t_split + res["mean_ms"]
    res["split_plus_mean_gbps_algorithmic"] = 24.0 * n / (total * 1e-3) * 1e-9
    res["frac_of_hbm_peak"] = res["split_plus_mean_gbps_algorithmic"] / hbm_peak
    res["hbm_peak_gbps"] = hbm_peak
    # the order-free one-pass flavour (vb200_block_reduce_stream): labels + atomics, no grouping
    cap = int(min(n, shape[0] * shape[1]))
    ids = torch.empty(cap, dtype=torch.int64, device="cuda")
    sout = torch.empty(cap, dtype=torch.float64, device="cuda")
    nb2 = ctypes.c_int64()
    t_stream = timed(lambda: _cabi.check(lib.vb200_block_reduce_stream(
        _cabi.ptr(de), _cabi.ptr(dn), n, _cabi.ptr(dce), ce.size, _cabi.ptr(dcn), cn.size, _cabi.ptr(dd), 1, None, 0, cap,
        _cabi.ptr(ids), _cabi.ptr(sout), ctypes.byref(nb2)), "stream"))
    res["order_free_mean_ms"] = t_stream
    res["order_free_gbps_algorithmic"] = 24.0 * n / (t_stream * 1e-3) * 1e-9
    res["order_free_frac_of_hbm_peak"] = res["order_free_gbps_algorithmic"] / hbm_peak
    res["order_free_same_blocks"] = bool(nb2.value == nblocks.value)
    res["order_free_max_rel_diff_vs_ordered"] = float(((sout[:nb2.value] - out).abs() / (out.abs() + 1e-300)).max().item()) if nb2.value == nblocks.value else None
    if n <= 400_000:  # the pandas + KD-tree reference-style host path beside it (oracle restatement)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import block_oracle as bo
        t0 = time.perf_counter()
        _, labels = bo.block_split((e, nn), shape=shape, region=(0, 5000, -5000, 0))
        bo.group_reduce(labels, d, np.mean)
        res["cpu_oracle_s"] = time.perf_counter() - t0
    print(json.dumps(res), flush=True)
    del de, dn, dd, out

for n, grid, maxdist in ((400_000, (2000, 2000), 5.0), (400_000, (10_000, 10_000), 5.0), (50_000, (10_000, 10_000), 15.0)):
    rng = np.random.RandomState(1)
    e, nn = rng.uniform(0, 5000, n), rng.uniform(-5000, 0, n)
    gx, gy = vb.grid_coordinates((0, 5000, -5000, 0), shape=grid)
    de, dn, qe, qn = dev(e), dev(nn), dev(gx.ravel()), dev(gy.ravel())
    mask = torch.empty(qe.numel(), dtype=torch.uint8, device="cuda")
    t = timed(lambda: _cabi.check(lib.vb200_distance_mask(_cabi.ptr(de), _cabi.ptr(dn), n, maxdist, _cabi.ptr(qe), _cabi.ptr(qn),
                                                          qe.numel(), _cabi.ptr(mask)), "mask"), reps=3)
    gbps = 17.0 * qe.numel() / (t * 1e-3) * 1e-9
    print(json.dumps({"case": "distance_mask", "data_points": n, "grid": list(grid), "maxdist": maxdist, "ms": t,
                      "nodes_per_s": qe.numel() / (t * 1e-3), "gbps_algorithmic": gbps, "frac_of_hbm_peak": gbps / hbm_peak,
                      "hbm_peak_gbps": hbm_peak, "inside_fraction": float(mask.float().mean().item())}), flush=True)
    del de, dn, qe, qn, mask
