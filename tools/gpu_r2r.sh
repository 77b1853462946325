#!/bin/bash
# final single-GPU validation of round 2: full -m gpu suite, smoke, default bench (with cpu_baseline), reference arm, configs 1 and 5
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/r2r_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2r_pytest_gpu.log | tail -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2r_smoke.log
timeout 900 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/r2r_bench.json; echo
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2r_reference.json 2> gpurun_out/r2r_reference.err; echo "reference rc=$?"; head -c 400 gpurun_out/r2r_reference.json; echo; tail -4 gpurun_out/r2r_reference.err
timeout 600 python tools/run_configs.py 1 5 > gpurun_out/r2r_configs.jsonl 2> gpurun_out/r2r_configs.err; echo "configs rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2r_configs.jsonl"):
    d = json.loads(l); d.pop("scores", None); d.pop("final_fit_info", None)
    if "fit_info" in d: d["fit_info"] = {k: d["fit_info"][k] for k in ("gram_ms", "factor_ms", "solve_ms", "diag_ms")}
    print(json.dumps(d))
PY
