#!/bin/bash
# N = 8: BASELINE config 5 (SplineCV jobs dealt to the ranks) and config 3 in full (400k rows, 10^8 grid nodes)
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 tools/run_config5_dist.py > gpurun_out/r2y_config5_${N}gpu.json 2> gpurun_out/r2y_config5_${N}gpu.err; echo "config5 rc=$?"; cut -c1-300 gpurun_out/r2y_config5_${N}gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --grid 3536 --steps 1 --warmup 1 --no-parity --no-e2e > gpurun_out/r2y_bench_config3_full_${N}gpu.json 2> gpurun_out/r2y_bench_config3_full_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2y_bench_config3_full_${N}gpu.json").read())
print(d["ms_per_step"], d["value"], d["fit_s"], d["predict_s"], d["predict_gpts_per_s"], d["fit_phases_ms"], d["config"]["grid_points"])
PY
