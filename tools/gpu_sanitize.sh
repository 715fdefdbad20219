#!/bin/bash
# compute-sanitizer over the hot-path cases of tools/sanitizer_cases.py (racecheck on the
# shared-memory BFS).  Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh r02a'
# Logs land in gpurun_out/sanitizer_<tool>_<tag>.log; the summaries are copied into profiles/.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
python tools/sanitizer_cases.py > $OUT/sanitizer_plain_$TAG.log 2>&1; echo "plain rc=$?"; cat $OUT/sanitizer_plain_$TAG.log
for tool in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  extra=""
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
  timeout ${CASE_TIMEOUT:-420} $CS --tool $tool $extra --print-limit 40 --log-file $OUT/sanitizer_${tool}_$TAG.log \
      python tools/sanitizer_cases.py ${CASES:-} > $OUT/sanitizer_${tool}_${TAG}.out 2>&1
  echo "$tool rc=$?"; tail -4 $OUT/sanitizer_${tool}_$TAG.log; tail -8 $OUT/sanitizer_${tool}_${TAG}.out
done
