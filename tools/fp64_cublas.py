"""cuBLAS FP64 yardsticks (library numbers, for context next to tools/fp64_peak.cu)."""
import json
import torch

def bench(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

out = {}
for n in (4096, 8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    ms = bench(lambda: torch.matmul(a, b))
    out["dgemm_%d_tflops" % n] = 2 * n**3 / ms * 1e-9
n = 16384
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
spd = a @ a.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
ms = bench(lambda: torch.linalg.cholesky(spd))
out["cusolver_potrf_%d_tflops" % n] = n**3 / 3 / ms * 1e-9
out["cusolver_potrf_%d_ms" % n] = ms
print(json.dumps(out))
