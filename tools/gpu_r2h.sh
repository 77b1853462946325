#!/bin/bash
# round 2, session 2, first pass: full -m gpu suite, default bench, ncu launch list, ncu --set full of the dynamic DMMA kernel,
# the elastic predict kernel and the elastic panel generator
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=10 > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2h_pytest_gpu.log | tail -30
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r2h_bench.json
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2h_launches.csv $CMD > gpurun_out/r2h_ncu_list.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vb_gemm_dynamic -s 3 -c 1 -o gpurun_out/r2h_gemm_dynamic $CMD > gpurun_out/r2h_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 900 python tools/run_configs.py 4 > gpurun_out/r2h_config4.jsonl 2> gpurun_out/r2h_config4.err; echo "config4 rc=$?"; cat gpurun_out/r2h_config4.jsonl | cut -c1-600
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vb_gram_panel -s 2 -c 1 -o gpurun_out/r2h_elastic_panel python tools/run_configs.py 4 > gpurun_out/r2h_ncu_elastic_panel.log 2>&1; echo "ncu elastic panel rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vb_predict2d -c 1 -o gpurun_out/r2h_elastic_predict python tools/run_configs.py 4 > gpurun_out/r2h_ncu_elastic_predict.log 2>&1; echo "ncu elastic predict rc=$?"
ls -la gpurun_out | grep r2h
