#!/bin/bash
# run on the GPU box: full -m gpu suite (no -x so that every failure is visible)
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider "$@" > gpurun_out/pytest_gpu.log 2>&1
rc=$?
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -40
exit $rc
