#!/bin/bash
N=${1:-4}
if [ "$3" = "bench" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r2n_bench_${N}gpu.json 2> gpurun_out/r2n_bench_${N}gpu.err; echo "bench N=$N rc=$?"
cat gpurun_out/r2n_bench_${N}gpu.json; grep -v "^\s*$\|OMP_NUM\|\*\*\*\*\|Under-determined\|warn(" gpurun_out/r2n_bench_${N}gpu.err | tail -5
fi
VB_PANEL_TILES=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/run_dist_pipeline.py 50000 > gpurun_out/r2n_pipeline_${N}gpu.jsonl 2> gpurun_out/r2n_pipeline_${N}gpu.err; echo "pipeline N=$N rc=$?"
cat gpurun_out/r2n_pipeline_${N}gpu.jsonl
