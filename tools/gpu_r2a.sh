#!/bin/bash
# round 2, first GPU pass: full -m gpu suite, short default bench, reference arm (all cores) with timing
rm -f gpurun_out/parity_observed.jsonl
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2a_pytest_gpu.log | tail -30
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.json
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2a_reference.json 2> gpurun_out/r2a_reference.err; echo "reference rc=$?"
tail -c 1200 gpurun_out/r2a_reference.json; tail -4 gpurun_out/r2a_reference.err
