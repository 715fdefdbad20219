#!/bin/bash
# End-of-round record: GPU tests + smoke + both bench arms, the ncu launch list of the short bench command, the debug-check suite.
#   gpurun --timeout 1500 -- 'bash tools/gpu_final.sh r02f'
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
bash tools/gpu_tests_bench.sh $TAG
BENCH_SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-latency --no-fhd --no-sequence"
$BENCH_SHORT > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    $BENCH_SHORT > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
bash tools/gpu_debug_checks.sh $TAG 2>&1 | tail -8
