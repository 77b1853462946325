#!/bin/bash
# N GPUs: panel-cyclic Cholesky schedules (serial / look-ahead / peer pushes, merged vs per-panel updates) at m = 50k,
# then the bench at N (config-3 workload + sharded-estimator parity block)
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/run_dist_pipeline.py 50000 > gpurun_out/r2j_pipeline_${N}gpu.jsonl 2> gpurun_out/r2j_pipeline_${N}gpu.err; echo "pipeline N=$N rc=$?"
cat gpurun_out/r2j_pipeline_${N}gpu.jsonl; tail -3 gpurun_out/r2j_pipeline_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r2j_bench_${N}gpu.json 2> gpurun_out/r2j_bench_${N}gpu.err; echo "bench N=$N rc=$?"
cat gpurun_out/r2j_bench_${N}gpu.json; tail -3 gpurun_out/r2j_bench_${N}gpu.err
