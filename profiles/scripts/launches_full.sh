#!/bin/bash
# launch list of the full bench workload (256 frames): per-kernel device times
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plainfull_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_full_$TAG.csv $CMD > gpurun_out/ncufull_$TAG.log 2>&1
tail -2 gpurun_out/ncufull_$TAG.log | cut -c1-300
