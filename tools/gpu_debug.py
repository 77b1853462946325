"""ad-hoc parity report printed on the GPU box (not a test)"""
import sys, os, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.simplefilter("ignore")
import verde_b200 as vb
import verde_oracle as vo
g = np.load(os.path.join(ROOT, "tests/golden/spline_small.npz"))
e, n, d, w = g["east"], g["north"], g["data"], g["weights"]
fc = (g["fc_east"], g["fc_north"])
sp = vb.Spline(damping=1e-5, force_coords=fc).fit((e, n), d, weights=w)
ref = g["force_fc"]
print("fc weighted: ferr", np.max(np.abs(sp.force_ - ref)) / np.max(np.abs(ref)), sp.fit_info_)
jac = vo.jacobian(e, n, fc[0], fc[1], 0.0)
G, b, c1, c2 = vo.normal_equations(jac, d, w)
s = vo.column_scale(jac)
A = G / s[:, None] / s[None, :]; A.flat[::61] += 1e-5
ev = np.linalg.eigvalsh(A); print("cond fc", ev[-1] / ev[0])
p = vb.least_squares(g["jac_fc"].copy(), d, w, damping=1e-5)
print("ls explicit: ferr", np.max(np.abs(p - ref)) / np.max(np.abs(ref)))
print("grid err", np.max(np.abs(sp.predict((g["grid_east"], g["grid_north"])) - g["grid_fc"])) / np.ptp(d))
for tag, damping, mindist, ww in (("d1em3", 1e-3, None, None), ("d1em6_md", 1e-6, 1e-5, None), ("d1em8", 1e-8, None, None), ("d1em4_w", 1e-4, None, w)):
    sp = vb.Spline(damping=damping, mindist=mindist).fit((e, n), d, weights=ww)
    ref = g["force_" + tag]
    print(tag, "cond %.2e" % g["cond_" + tag], "ferr %.2e" % (np.max(np.abs(sp.force_ - ref)) / np.max(np.abs(ref))),
          "gerr %.2e" % (np.max(np.abs(sp.predict((g["grid_east"], g["grid_north"])) - g["grid_" + tag])) / np.ptp(d)),
          "diag", sp.fit_info_["diag_min"], sp.fit_info_["diag_max"])
c = np.load(os.path.join(ROOT, "tests/golden/cfg1_5k.npz"))
tab = vb.CheckerBoard().scatter(size=5000, random_state=0)
e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
full = vb.grid_coordinates((e.min(), e.max(), n.min(), n.max()), shape=(500, 500))
sub = (full[0][::10, ::10], full[1][::10, ::10])
for tag, damping, mindist in (("wellcond", 1e-1, None), ("cfg2like", 1e-6, 1e-5), ("cfg1", 1e-8, None)):
    sp = vb.Spline(damping=damping, mindist=mindist).fit((e, n), d)
    print(tag, "ferr %.2e" % (np.max(np.abs(sp.force_ - c["force_" + tag])) / np.max(np.abs(c["force_" + tag]))),
          "gerr %.2e" % (np.max(np.abs(sp.predict(sub) - c["grid_sub_" + tag])) / np.ptp(d)),
          "derr %.2e" % (np.max(np.abs(sp.predict((e[:500], n[:500])) - c["pred_data_" + tag])) / np.ptp(d)),
          {k: (round(v, 3) if isinstance(v, float) else v) for k, v in sp.fit_info_.items()})
