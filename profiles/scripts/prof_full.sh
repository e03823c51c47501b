#!/bin/bash
# one --set full capture of the two dominant kernels at the full bench workload (256 frames), raw + source pages
# exported, and profiles/k1_traffic.json rewritten from it (dram bytes of one K1 launch + hash of the kernel sources)
TAG=${1:-r02}
CMD="python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plainfull2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_fused_pair|k_fused_stream|k_entropy_fused" -s 8 -c 2 -f -o gpurun_out/proffull_$TAG $CMD > gpurun_out/ncufull2_$TAG.log 2>&1
tail -2 gpurun_out/ncufull2_$TAG.log | cut -c1-200
ncu -i gpurun_out/proffull_$TAG.ncu-rep --page raw --csv > gpurun_out/proffull_${TAG}_raw.csv
ncu -i gpurun_out/proffull_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/proffull_${TAG}_sass.csv
python profiles/scripts/ncu_summary.py gpurun_out/proffull_${TAG}_raw.csv > gpurun_out/ncu_full_workload_summary_$TAG.txt
python - "$TAG" <<'PY'
import csv, json, sys
sys.path.insert(0, ".")
import bench
rows = list(csv.reader(open("gpurun_out/proffull_%s_raw.csv" % sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    v = float(r[ix[name]].replace(",", ""))
    u = units[ix[name]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
for r in rows[2:]:
    if "k_fused_pair" in r[ix["Kernel Name"]] or "k_fused_stream" in r[ix["Kernel Name"]]:
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        json.dump({"bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                   "k1_source_sha": bench.k1_source_sha(),
                   "source": "ncu --set full capture of %s at the bench workload (256 x 1080p 4:2:0), dram__bytes_read.sum + "
                             "dram__bytes_write.sum; summary in profiles/%s/ncu_full_workload_summary.txt (profiles/scripts/prof_full.sh)"
                             % (r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), sys.argv[1])},
                  open("gpurun_out/k1_traffic.json", "w"))
        print(open("gpurun_out/k1_traffic.json").read())
        break
PY
