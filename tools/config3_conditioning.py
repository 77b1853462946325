"""
Which dampings make BASELINE config 3's normal matrix numerically positive definite in FP64?  One GPU.
Builds the N>1 bench workload (first n of config 3's 400 000 CheckerBoard points, forces at the BlockReduce centres of
all 400 000, optional weights RandomState(1).uniform(0.1, 10)) for n = 100k / 200k / 400k (the N = 2 / 4 / 8 bench
sizes; rows are accumulated incrementally into ONE Gram matrix) and factors it for a list of dampings side by side
(vb200_gram_solve_multi): prints the failing pivot (0 = factored) per damping, lambda_max and the smallest pivot.
    python tools/config3_conditioning.py
"""
import ctypes, json, os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import torch
import verde_b200 as vb
from verde_b200 import _cabi, kernels
import bench

lib = _cabi.require_device()
east, north, data, weights = bench.config3_points()
(fe, fn), _ = vb.BlockReduce(np.mean, shape=bench.CONFIG3_BLOCKS, center_coordinates=True).filter((east, north), data)
m = fe.size
dampings = [float(a) for a in sys.argv[1:]] or [1e-6, 4e-6, 1e-5, 4e-5, 1e-4]
for weighted in (True, False):
    system = kernels.GramSystem(m, peer_exchange=False)
    done = 0
    for n in (100_000, 200_000, 400_000):
        w = weights[done:n] if weighted else None
        system.accumulate(0, east[done:n], north[done:n], data[done:n], w, fe, fn, bench.MINDIST, accumulate=done > 0)
        done = n
        damp = np.ascontiguousarray(dampings)
        params = np.empty((damp.size, m))
        pivots = np.zeros(damp.size, dtype=np.int32)
        info = _cabi.FitInfo()
        torch.cuda.synchronize()
        rc = lib.vb200_gram_solve_multi(_cabi.ptr(system.gram), _cabi.ptr(system.stats), m, n, _cabi.ptr(damp), damp.size,
                                        _cabi.ptr(params), _cabi.ptr(pivots), ctypes.byref(info))
        sum_w = float(np.sum(w)) if weighted else float(n)
        # one more, single, solve with the damping scaled like the total weight (config 2's relative damping)
        scaled = 1e-6 * float(np.sum(weights[:n]) if weighted else n) / 50_000.0
        out = {"rows": n, "forces": int(m), "weighted": weighted, "sum_w": float(np.sum(weights[:n])) if weighted else float(n),
               "dampings": dampings, "failed_pivot": [int(p) for p in pivots], "rc": int(rc),
               "factor_ms": info.factor_ms}
        info2 = _cabi.FitInfo()
        f = np.empty(m)
        rc2 = lib.vb200_gram_solve(_cabi.ptr(system.gram), _cabi.ptr(system.stats), m, n, scaled, 1, _cabi.ptr(f), ctypes.byref(info2))
        out["scaled_damping"] = {"damping": scaled, "rc": int(rc2), "diag_min": info2.diag_min, "diag_max": info2.diag_max,
                                 "lambda_max": info2.lambda_max, "cond_est": info2.cond_est, "cond_upper": info2.cond_upper}
        print(json.dumps(out), flush=True)
    del system
    torch.cuda.empty_cache()
