#!/bin/bash
# damping=None (Jacobi SVD) parity + timing
rm -f gpurun_out/parity_observed.jsonl
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "damping_none or default_constructors or weights_without" --durations=8 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2b_pytest.log
cat gpurun_out/parity_observed.jsonl
