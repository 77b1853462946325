#!/bin/bash
# The validation sequence of a round's final state (what tools/gpu_r2r.sh, gpu_r2m.sh and gpu_r2y.sh ran in round 2):
#   1 GPU :  bash tools/gpu_final.sh
#   N GPUs:  bash tools/gpu_final.sh N        (gpurun --gpus N)
N=${1:-1}
if [ "$N" = "1" ]; then
    bash tools/gpu_r2r.sh
else
    bash tools/gpu_r2m.sh $N 4
fi
