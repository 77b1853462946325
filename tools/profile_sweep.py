"""One (fold, mindist) job of BASELINE config 5 for profiling: Gram matrix of 16 000 points, then the six dampings
factored side by side (vb200_gram_solve_multi).   python tools/profile_sweep.py"""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import verde_b200 as vb
from verde_b200 import kernels

tab = vb.CheckerBoard().scatter(size=16000, random_state=0)
e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
if len(sys.argv) > 1:
    assert vb._cabi.require_device().vb200_set_option(b"trsm_tiles", float(sys.argv[1])) == 0
system = kernels.GramSystem(e.size).accumulate(0, e, n, d, None, e, n, 0.0)
forces = system.solve_multi(e.size, [1e-8, 1e-6, 1e-4, 1e-3, 1e-2, 1e-1])
print({k: getattr(system.info, k) for k in ("gram_ms", "factor_ms", "solve_ms", "gemm_ms", "gemm_flops", "gemm_launches")})
