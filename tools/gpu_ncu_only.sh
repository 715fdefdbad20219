#!/bin/bash
# The two ncu passes of tools/gpu_round.sh without the test / bench legs (B200_PROFILING.md recipe):
#   gpurun --timeout 600 -- 'bash tools/gpu_ncu_only.sh r01h'
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
BENCH_SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-latency"
$BENCH_SHORT > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    $BENCH_SHORT > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
$BENCH_SHORT > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'cell_stats|region_grow|label|edge_mask|refine|depth' -s 16 -c 6 \
    -f -o $OUT/prof_$TAG $BENCH_SHORT > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
