#!/bin/bash
# compute-sanitizer over the smoke test and one cfg2-shaped step (run under gpurun; summaries land in gpurun_out/).
# the kernels use global / shared atomics, mbarriers and async copies -> memcheck + racecheck (+ synccheck).
set -u
out=gpurun_out
mkdir -p $out
for tool in memcheck racecheck synccheck; do
  for target in "python __graft_entry__.py --smoke" "python tools/sanitize_step.py"; do
    name=$(echo $target | tr ' /.' '___')
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 $target > $out/sanitizer_${tool}_${name}.log 2>&1
    echo "== $tool :: $target :: exit $?" >> $out/sanitizer_summary.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" $out/sanitizer_${tool}_${name}.log | sort | uniq -c | head -20 >> $out/sanitizer_summary.txt
  done
done
cat $out/sanitizer_summary.txt
