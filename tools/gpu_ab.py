"""A/B of tunables on the GPU box: python tools/gpu_ab.py <n> key=v1,v2,... [key2=...]"""
import ctypes, itertools, os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import verde_b200 as vb
from verde_b200 import _cabi
import torch

n = int(sys.argv[1])
sweeps = [(a.split("=")[0], [float(v) for v in a.split("=")[1].split(",")]) for a in sys.argv[2:]]
lib = _cabi.require_device()
rng = np.random.RandomState(0)
e = rng.uniform(0, 5000, n); nn = rng.uniform(-5000, 0, n)
d = 1000.0 * np.sin((2 * np.pi / 2500.0) * e) * np.cos((2 * np.pi / 2500.0) * nn)
dev = torch.device("cuda", 0)
de, dn, dd = (torch.from_numpy(x).to(dev) for x in (e, nn, d))
out = torch.empty(n, dtype=torch.float64, device=dev)
ref = None
print("peak dmma", lib.vb200_measure_fp64_peak_tflops(0))
for combo in itertools.product(*[v for _, v in sweeps]):
    for (k, _), v in zip(sweeps, combo):
        assert lib.vb200_set_option(k.encode(), v) == 0, k
    info = _cabi.FitInfo()
    best = None
    for rep in range(2):
        t0 = time.perf_counter()
        rc = lib.vb200_spline_fit(_cabi.ptr(de), _cabi.ptr(dn), _cabi.ptr(dd), None, n, _cabi.ptr(de), _cabi.ptr(dn), n,
                                  1e-5, 1e-6, _cabi.ptr(out), ctypes.byref(info))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rc != 0:
            print(combo, "rc", rc, lib.vb200_last_error()); break
        i = info.as_dict()
        best = (dt, i)
    if best is None: continue
    f = out.cpu().numpy()
    if ref is None: ref = f.copy()
    dt, i = best
    print(dict(zip([k for k, _ in sweeps], combo)), "wall %.3f s gram %.1f factor %.1f solve %.1f diag %.1f (inv its %d, cond_est %.3e <= %.3e, bwd %.1e) gemm %.1f ms -> %.2f TF; max|f-f0|/max|f0| = %.2e" % (
        dt, i["gram_ms"], i["factor_ms"], i["solve_ms"], i["diag_ms"], i["inverse_its"], i["cond_est"], i["cond_upper"], i["backward_err"],
        i["gemm_ms"], i["gemm_flops"] / i["gemm_ms"] * 1e-9,
        np.max(np.abs(f - ref)) / np.max(np.abs(ref))), flush=True)
