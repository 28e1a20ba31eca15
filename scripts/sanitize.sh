#!/bin/bash
# compute-sanitizer over the small-shape workload of scripts/sanitize_targets.py; logs -> gpurun_out/sanitizer_<tool>_<part>.log
# usage (GPU box): bash scripts/sanitize.sh "memcheck synccheck racecheck" "decode seq train" [per-run timeout seconds]
TOOLS=${1:-"memcheck synccheck racecheck"}
PARTS=${2:-"decode seq train"}
TMO=${3:-420}
mkdir -p gpurun_out
python scripts/sanitize_targets.py all > gpurun_out/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitizer_plain.log; exit 1; }
for tool in $TOOLS; do
  for part in $PARTS; do
    log=gpurun_out/sanitizer_${tool}_${part}.log
    extra=""
    [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
    timeout $TMO compute-sanitizer --tool $tool $extra --print-limit 40 --error-exitcode 9 python scripts/sanitize_targets.py $part > $log 2>&1
    rc=$?
    echo "$tool $part rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
  done
done
