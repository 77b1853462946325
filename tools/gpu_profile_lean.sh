#!/bin/bash
# one full ncu capture of the dominant Gram launch and of the predict launch (no launch list)
tag=${1:-r01c}
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:vb_gemm_persistent -s 3 -c 1 -o gpurun_out/${tag}_gemm $CMD > gpurun_out/${tag}_ncu_gemm.log 2>&1
$CMD > gpurun_out/${tag}_plain2.json 2> gpurun_out/${tag}_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:vb_predict_kernel -s 1 -c 1 -o gpurun_out/${tag}_predict $CMD > gpurun_out/${tag}_ncu_predict.log 2>&1
ls -la gpurun_out | grep ${tag}
