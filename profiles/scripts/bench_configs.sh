#!/bin/bash
# bench.py over every BASELINE.json config (+ the two 8K stress distributions of SURVEY 8d), one JSON line each
TAG=${1:-r01}
OUT=gpurun_out/configs_$TAG.jsonl
mkdir -p gpurun_out
: > $OUT
for cfg in 1 2 3a 3b 4; do
    python bench.py --config $cfg --steps 10 --warmup 3 2>gpurun_out/configs_$TAG.$cfg.err | tail -1 >> $OUT
done
for kind in noise grey; do
    python bench.py --config 4 --kind $kind --steps 10 --warmup 3 --skip-cpu-baseline 2>gpurun_out/configs_$TAG.4$kind.err | tail -1 >> $OUT
done
python - <<PY
import json
for l in open("$OUT"):
    try:
        d = json.loads(l)
    except Exception:
        print("bad line:", l[:200]); continue
    e = d.get("e2e") or {}
    c = d.get("cpu_baseline") or {}
    r = d["roofline"]
    print("%-95s | %9.0f Mpx/s | %7.3f ms | K1 %6.3f ms %5.1f%% hbm | e2e %8.0f | cpu %7.0f (%s thr) | %.2f bpp | %s MHz %s" % (
        d["config"]["workload"][:95], d["value"], d["ms_per_step"], r["ms_per_launch"], 100 * r["frac"], e.get("value", 0),
        c.get("value", 0), c.get("cores"), d["bits_per_pixel"], d["clocks"].get("sm_mhz"), d["clocks"].get("reasons")))
PY
