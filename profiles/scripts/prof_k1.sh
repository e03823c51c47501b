#!/bin/bash
# ncu recipe of /opt/skills/guides/B200_PROFILING.md: launch list, then one --set full capture.
# usage (under gpurun): bash profiles/scripts/prof_k1.sh <tag> [kernel-regex]
TAG=${1:-r01}
REGEX=${2:-'k_fused_tiled|k_bitpack|k_stuff_write|k_block_bits'}
SKIP=${3:-8}
COUNT=${4:-4}
CMD="python bench.py --frames 32 --steps 2 --warmup 3 --skip-e2e --skip-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $COUNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log | cut -c1-600; tail -3 gpurun_out/ncu1_$TAG.log; tail -3 gpurun_out/ncu2_$TAG.log; ls -la gpurun_out
