#!/bin/bash
# compute-sanitizer over smoke() and the entropy kernel's exposed cases (profiles/scripts/sanitize_target.py);
# one log per tool under gpurun_out/ (copied to profiles/<round>/ by hand).  SURVEY section 5 "race detection / sanitizers".
TAG=${1:-r02}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python profiles/scripts/sanitize_target.py \
      > gpurun_out/sanitizer_${tool}_$TAG.txt 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok$|ok:" gpurun_out/sanitizer_${tool}_$TAG.txt | tail -8
done
