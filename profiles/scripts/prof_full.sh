#!/bin/bash
# one --set full capture of the two dominant kernels at the full bench workload (256 frames)
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plainfull2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_fused_tiled|k_entropy_fused" -s 8 -c 2 -f -o gpurun_out/proffull_$TAG $CMD > gpurun_out/ncufull2_$TAG.log 2>&1
tail -2 gpurun_out/ncufull2_$TAG.log | cut -c1-200
