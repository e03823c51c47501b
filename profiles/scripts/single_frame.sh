#!/bin/bash
# BASELINE configs[1] (one 1080p frame): eager and replayed from a CUDA graph, 400 steps each
TAG=${1:-r02}
mkdir -p gpurun_out
python bench.py --config 1 --steps 400 --warmup 40 --skip-cpu-baseline > gpurun_out/bench_${TAG}_cfg1.json 2> gpurun_out/bench_${TAG}_cfg1.err
python bench.py --config 1 --steps 400 --warmup 40 --skip-cpu-baseline --graph > gpurun_out/bench_${TAG}_cfg1_graph.json 2> gpurun_out/bench_${TAG}_cfg1_graph.err
python - <<PY
import json
for n in ("cfg1", "cfg1_graph"):
    d = json.loads(open("gpurun_out/bench_${TAG}_%s.json" % n).read().strip().split("\n")[-1])
    print(n, "%.2f us per frame" % (1000 * d["ms_per_step"]), d.get("stage_ms"))
PY
