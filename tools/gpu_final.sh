#!/bin/bash
# End-of-round evidence on one B200: the driver's default bench commands, configs 1/4/5, the ncu launch
# list of the bench command and one full capture of the new block / mask kernels.
# usage: tools/gpu_final.sh <tag>
tag=${1:-r01d}
python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"
python tools/run_configs.py 1 4 5 > gpurun_out/${tag}_configs.jsonl 2> gpurun_out/${tag}_configs.err; echo "configs rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "launch list rc=$?"
python tools/run_blocks.py > gpurun_out/${tag}_blocks.jsonl 2> gpurun_out/${tag}_blocks.err; echo "blocks rc=$?"
# case 2 of run_blocks.py (20M points, 224x224 blocks): its first scatter launch is the 73rd; the mask's second grid is its 5th launch
ncu --set full --clock-control none --import-source on -k regex:vb_rsort_scatter -s 72 -c 2 -o gpurun_out/${tag}_rsort python tools/run_blocks.py > gpurun_out/${tag}_ncu_rsort.log 2>&1; echo "ncu rsort rc=$?"
ncu --set full --clock-control none --import-source on -k regex:vb_distance_mask_kernel -s 4 -c 1 -o gpurun_out/${tag}_mask python tools/run_blocks.py > gpurun_out/${tag}_ncu_mask.log 2>&1; echo "ncu mask rc=$?"
ls -la gpurun_out | grep ${tag}
