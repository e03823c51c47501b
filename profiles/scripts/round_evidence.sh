#!/bin/bash
# Everything the round's numbers come from, in one gpurun call (1 GPU):
#   1. pytest -m gpu           2. bench.py (default)      3. bench.py --impl reference
#   4. launch list (ncu gpu__time_duration, after the same command ran clean without ncu)
# The --set full capture is a separate call (profiles/scripts/prof_full.sh): one profiler per call.
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 400 gpurun_out/bench_ref_$TAG.json
bash profiles/scripts/launches_full.sh $TAG
python profiles/scripts/launch_summary.py gpurun_out/launches_full_$TAG.csv
