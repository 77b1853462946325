"""
Times BASELINE.json configs 1, 4 and 5 through the estimator API on one GPU and prints one
JSON object per config (config 2 is bench.py's default; config 3 is bench.py at N = 8).
    python tools/run_configs.py [1] [4] [5]
"""
import json, os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import verde_b200 as vb
import torch

which = [int(a) for a in sys.argv[1:]] or [1, 4, 5]


def timed(fn, reps=2):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out


if 1 in which:
    tab = vb.CheckerBoard().scatter(size=5000, random_state=0)
    coords, d = (tab.easting.values, tab.northing.values), tab.scalars.values
    sp = vb.Spline(damping=1e-8)
    t_fit, _ = timed(lambda: sp.fit(coords, d))
    grid = vb.grid_coordinates(sp.region_, shape=(500, 500))
    t_pred, pred = timed(lambda: sp.predict(grid))
    print(json.dumps({"config": 1, "fit_s": t_fit, "predict_s": t_pred, "predict_gpairs_per_s": 250000 * 5000 / t_pred * 1e-9,
                      "fit_info": sp.fit_info_, "max_abs_resid_at_data": float(np.max(np.abs(sp.predict(coords) - d)))}), flush=True)

if 4 in which:
    n = 20000
    e, nn = vb.scatter_points((0, 500e3, 0, 500e3), n, random_state=0)
    ue = 10 * np.sin(2 * np.pi * e / 3e5) * np.cos(2 * np.pi * nn / 4.5e5) + 3
    un = -7 * np.cos(2 * np.pi * e / 4.5e5) * np.sin(2 * np.pi * nn / 3e5) + 1
    vs = vb.VectorSpline2D(poisson=0.5, mindist=10e3, damping=1e-3)
    t_fit, _ = timed(lambda: vs.fit((e, nn), (ue, un)), reps=2)
    grid = vb.grid_coordinates(vs.region_, shape=(1000, 1000))
    t_pred, pred = timed(lambda: vs.predict(grid))
    pe, pn = vs.predict((e, nn))
    info = vs.fit_info_
    flops = 2 * n * (2 * n) * (2 * n + 1) + (2 * n) ** 3 / 3
    print(json.dumps({"config": 4, "stations": n, "system": [2 * n, 2 * n], "damping": 1e-3, "fit_s": t_fit,
                      "fit_tflops_algorithmic": flops / t_fit * 1e-12, "predict_s_1Mpts": t_pred,
                      "predict_gpairs_per_s": 1e6 * n / t_pred * 1e-9, "fit_info": info,
                      "rms_resid_east": float(np.sqrt(np.mean((pe - ue) ** 2))), "rms_resid_north": float(np.sqrt(np.mean((pn - un) ** 2)))}), flush=True)

if 5 in which:
    tab = vb.CheckerBoard().scatter(size=20000, random_state=0)
    coords, d = (tab.easting.values, tab.northing.values), tab.scalars.values
    dampings = [1e-8, 1e-6, 1e-4, 1e-3, 1e-2, 1e-1]
    mindists = [0, 1e-5, 1e-3, 1e-1]
    cv = vb.SplineCV(dampings=dampings, mindists=mindists)
    t_fit, _ = timed(lambda: cv.fit(coords, d), reps=2)
    print(json.dumps({"config": 5, "points": 20000, "combos": 24, "folds": 5, "fit_s": t_fit, "scores": list(map(float, cv.scores_)),
                      "best": [float(cv.mindist_), float(cv.damping_)], "sweep_info": cv.sweep_info_,
                      "final_fit_info": cv.spline_.fit_info_}), flush=True)
