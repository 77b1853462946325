#!/bin/bash
# blocked diagonal solve in the sweeps, rsqrt potrf, capped inverse iteration: correctness, then timings at config 2 / 5 / 1 sizes
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2i_pytest_gpu.log | tail -30
timeout 600 python tools/gpu_ab.py 5000 cond_inverse_its=0 > gpurun_out/r2i_ab5k.log 2>&1; cat gpurun_out/r2i_ab5k.log
timeout 600 python tools/gpu_ab.py 16000 cond_inverse_its=0 > gpurun_out/r2i_ab16k.log 2>&1; cat gpurun_out/r2i_ab16k.log
timeout 900 python tools/gpu_ab.py 50000 cond_inverse_its=0,8 > gpurun_out/r2i_ab50k.log 2>&1; cat gpurun_out/r2i_ab50k.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
head -c 1500 gpurun_out/r2i_bench.json
timeout 600 python tools/run_configs.py 1 5 > gpurun_out/r2i_configs.jsonl 2> gpurun_out/r2i_configs.err; echo "configs rc=$?"; cut -c1-700 gpurun_out/r2i_configs.jsonl
