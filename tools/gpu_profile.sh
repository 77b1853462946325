#!/bin/bash
# ncu launch list + one full capture of the dominant kernel, on the bench command.
# usage: tools/gpu_profile.sh <tag>
tag=${1:-r01}
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1
$CMD > gpurun_out/${tag}_plain2.json 2> gpurun_out/${tag}_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:vb_gemm_persistent -s 3 -c 2 -o gpurun_out/${tag}_gemm $CMD > gpurun_out/${tag}_ncu_full.log 2>&1
$CMD > gpurun_out/${tag}_plain3.json 2> gpurun_out/${tag}_plain3.err &&
ncu --set full --clock-control none --import-source on -k regex:vb_predict_kernel -s 1 -c 1 -o gpurun_out/${tag}_predict $CMD > gpurun_out/${tag}_ncu_predict.log 2>&1
ls -la gpurun_out | tail -12
