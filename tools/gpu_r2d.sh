#!/bin/bash
# dynamic tile queue + look-ahead Cholesky: correctness, then A/B at config 2 / 5-ish / 1 sizes
timeout 600 python -m pytest tests/test_gpu_variants.py -q -p no:cacheprovider -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2d_pytest.log
timeout 900 python tools/gpu_ab.py 50000 gemm_variant=3,4 > gpurun_out/r2d_ab50k_variant.log 2>&1; cat gpurun_out/r2d_ab50k_variant.log
timeout 900 python tools/gpu_ab.py 50000 gemm_variant=4 chol_lookahead=8,16,24,32 > gpurun_out/r2d_ab50k_ahead.log 2>&1; cat gpurun_out/r2d_ab50k_ahead.log
timeout 600 python tools/gpu_ab.py 16000 gemm_variant=3,4 chol_lookahead=0,16 > gpurun_out/r2d_ab16k.log 2>&1; cat gpurun_out/r2d_ab16k.log
timeout 600 python tools/gpu_ab.py 5000 gemm_variant=3,4 chol_lookahead=0,8 chol_lookahead_min_rows=16 > gpurun_out/r2d_ab5k.log 2>&1; cat gpurun_out/r2d_ab5k.log
