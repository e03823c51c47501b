#!/bin/bash
# ncu --set full of k_entropy_fused for several library builds (ab_entropy.py launches them)
V=jpeg_encoder_gsoc2017_ishan-bansal_b200/variants
for n in "$@"; do
  ncu --set full --clock-control none -k regex:k_entropy_fused -s 4 -c 1 -f -o gpurun_out/ab_$n python profiles/scripts/ab_entropy.py $V/libb2j_$n.so > gpurun_out/ab_$n.log 2>&1
  ncu -i gpurun_out/ab_$n.ncu-rep --page raw --csv > gpurun_out/ab_${n}_raw.csv
  python profiles/scripts/ncu_summary.py gpurun_out/ab_${n}_raw.csv > gpurun_out/ab_${n}_summary.txt
  rm -f gpurun_out/ab_$n.ncu-rep
done
