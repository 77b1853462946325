"""
Parity at the HEADLINE size (BASELINE config 2: 50 000 points, forces at the data, damping 1e-6,
mindist 1e-5): the CPU oracle (bit-exact restatement of the reference, tests/test_oracle.py) is run on
the GPU box's host cores (needs ~80 GB of RAM and ~10 min) and compared with the B200 path on the
same inputs.  Prints one JSON object.   python tools/parity_full_config2.py [n [damping]]
"""
import faulthandler, json, os, sys, time, warnings
faulthandler.enable()
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.simplefilter("ignore")
import scipy.linalg
import verde_oracle as vo
import verde_b200 as vb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
damping = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-6
mindist = 1e-5
tab = vb.CheckerBoard().scatter(size=n, random_state=0)
e, nn, d = tab.easting.values, tab.northing.values, tab.scalars.values
out = {"n": n, "damping": damping, "mindist": mindist, "host_cores": os.cpu_count()}

t0 = time.perf_counter()
sp = vb.Spline(damping=damping, mindist=mindist).fit((e, nn), d)
out["gpu_fit_s"] = time.perf_counter() - t0
out["gpu_fit_info"] = sp.fit_info_

# --- oracle, memory-lean but the same operations as oracle/verde_oracle.py::least_squares ---
t0 = time.perf_counter()
jac = vo.jacobian(e, nn, e, nn, mindist)
out["cpu_jacobian_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
mean = jac.sum(axis=0) / n
var = np.zeros(n); corr = np.zeros(n)
for a in range(0, n, 2000):  # corrected two-pass variance in row slabs (sklearn's formula)
    tmp = jac[a:a + 2000] - mean
    corr += tmp.sum(axis=0); var += (tmp * tmp).sum(axis=0)
var = (var - corr**2 / n) / n
scale = np.sqrt(var)
for a in range(0, n, 2000):
    jac[a:a + 2000] /= scale
out["cpu_scaler_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
gram = jac.T @ jac
rhs = jac.T @ d
out["cpu_gram_s"] = time.perf_counter() - t0
lam_max = None
v = np.random.RandomState(0).normal(size=n)
for _ in range(30):  # power iteration for lambda_max of the scaled normal matrix
    v = gram @ v; lam_max = np.linalg.norm(v); v /= lam_max
out["lambda_max"] = float(lam_max); out["cond_estimate"] = float(lam_max / damping)
t0 = time.perf_counter()
gram.flat[:: n + 1] += damping
coef = scipy.linalg.solve(gram, rhs, assume_a="pos", overwrite_a=True, check_finite=False)
force_ref = coef / scale
out["cpu_solve_s"] = time.perf_counter() - t0
del gram

gx, gy = vb.grid_coordinates(sp.region_, shape=(200, 200))
t0 = time.perf_counter()
grid_ref = vo.predict(gx, gy, e, nn, mindist, force_ref)
out["cpu_predict_40k_pts_s"] = time.perf_counter() - t0
grid_gpu = sp.predict((gx, gy))
grid_gpu_refforce = np.empty(gx.size)
vb.kernels.predict_b200(gx.ravel(), gy.ravel(), e, nn, mindist, force_ref, grid_gpu_refforce)
rng = np.ptp(d)
out["force_gap_rel_max"] = float(np.max(np.abs(sp.force_ - force_ref)) / np.max(np.abs(force_ref)))
out["grid_gap_rel_range"] = float(np.max(np.abs(grid_gpu - grid_ref)) / rng)
out["predict_kernel_gap_same_forces_rel_range"] = float(np.max(np.abs(grid_gpu_refforce.reshape(gx.shape) - grid_ref)) / rng)
# which of the two solutions satisfies the reference's own equations better?  residual of
# (Js^T Js + alpha I) c = Js^T d for both, with the CPU's scaled Jacobian (jac is Js now)
for name, f in (("cpu", force_ref), ("gpu", sp.force_)):
    c = f * scale
    r = jac.T @ (d - jac @ c) - damping * c
    out["normal_eq_residual_rel_" + name] = float(np.linalg.norm(r) / (lam_max * np.linalg.norm(c) + np.linalg.norm(rhs)))
    out["data_misfit_rms_" + name] = float(np.sqrt(np.mean((jac @ c - d) ** 2)))
print(json.dumps(out), flush=True)
