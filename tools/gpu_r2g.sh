#!/bin/bash
N=${1:-2}
shift
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/peer_probe.py > gpurun_out/r2g_peer_probe_${N}gpu.jsonl 2> gpurun_out/r2g_peer_probe_${N}gpu.err; echo "probe rc=$?"
cat gpurun_out/r2g_peer_probe_${N}gpu.jsonl; tail -3 gpurun_out/r2g_peer_probe_${N}gpu.err
bash tools/gpu_r2f.sh $N "$@"
