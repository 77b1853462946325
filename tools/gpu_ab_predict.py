"""A/B of predict tunables: python tools/gpu_ab_predict.py <npts> <m> key=v1,v2 ..."""
import itertools, os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verde_b200 import _cabi
import torch
npts, m = int(sys.argv[1]), int(sys.argv[2])
sweeps = [(a.split("=")[0], [float(v) for v in a.split("=")[1].split(",")]) for a in sys.argv[3:]]
lib = _cabi.require_device()
rng = np.random.RandomState(0)
dev = torch.device("cuda", 0)
t = lambda x: torch.from_numpy(x).to(dev)
e, n = t(rng.uniform(0, 5000, npts)), t(rng.uniform(-5000, 0, npts))
fe, fn, f = t(rng.uniform(0, 5000, m)), t(rng.uniform(-5000, 0, m)), t(rng.normal(size=m))
out = torch.empty(npts, dtype=torch.float64, device=dev)
ref = None
for mindist in (1e-5, 0.0):
    for combo in itertools.product(*[v for _, v in sweeps]):
        for (k, _), v in zip(sweeps, combo):
            assert lib.vb200_set_option(k.encode(), v) == 0, k
        best = 1e30
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.vb200_predict(_cabi.ptr(e), _cabi.ptr(n), npts, _cabi.ptr(fe), _cabi.ptr(fn), _cabi.ptr(f), m, mindist, _cabi.ptr(out))
            e1.record(); torch.cuda.synchronize()
            assert rc == 0
            best = min(best, e0.elapsed_time(e1))
        o = out.cpu().numpy()
        if ref is None or ref[0] != mindist: ref = (mindist, o.copy())
        print("mindist", mindist, dict(zip([k for k, _ in sweeps], combo)), "%.2f ms  %.1f Gpairs/s  maxdiff %.2e" % (
            best, npts * m / best * 1e-6, np.max(np.abs(o - ref[1])) / np.max(np.abs(ref[1]))), flush=True)
