#!/bin/bash
# N=2: bench (config-3 weak-scaled workload + sharded-estimator parity block) and the reference arm under torchrun
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2c_bench_${N}gpu.json 2> gpurun_out/r2c_bench_${N}gpu.err; echo "bench N=$N rc=$?"
tail -c 3000 gpurun_out/r2c_bench_${N}gpu.json; tail -5 gpurun_out/r2c_bench_${N}gpu.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2c_reference_${N}gpu.json 2> gpurun_out/r2c_reference_${N}gpu.err; echo "reference N=$N rc=$?"
tail -c 900 gpurun_out/r2c_reference_${N}gpu.json; tail -4 gpurun_out/r2c_reference_${N}gpu.err
