#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_variants.py -q -p no:cacheprovider > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2e_pytest.log
timeout 1200 python tools/gpu_ab.py 50000 gemm_variant=4 chol_lookahead=24,32,48 chol_lookahead_min_rows=48,64 > gpurun_out/r2e_ab50k_ahead.log 2>&1; cat gpurun_out/r2e_ab50k_ahead.log
timeout 600 python tools/gpu_ab.py 16000 gemm_variant=4 chol_lookahead=0,32 chol_lookahead_min_rows=48,64 > gpurun_out/r2e_ab16k.log 2>&1; cat gpurun_out/r2e_ab16k.log
