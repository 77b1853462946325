#!/bin/bash
N=${1:-2}
shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/run_dist_pipeline.py "$@" > gpurun_out/r2f_pipeline_${N}gpu.jsonl 2> gpurun_out/r2f_pipeline_${N}gpu.err; echo "pipeline N=$N rc=$?"
cat gpurun_out/r2f_pipeline_${N}gpu.jsonl; tail -5 gpurun_out/r2f_pipeline_${N}gpu.err
