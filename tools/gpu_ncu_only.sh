#!/bin/bash
# The ncu passes behind profiles/ (B200_PROFILING.md recipe), each only after the same command has exited 0 without ncu:
#   gpurun --timeout 900 -- 'bash tools/gpu_ncu_only.sh r02d'
#   1. launch list (gpu__time_duration.sum) of the short bench command;
#   2. ncu --set full of the stage kernels in that command;
#   3. the same with stage 3 unfused (DPX_FUSE_LABELING=0): the labeling kernel writing the whole batch alone;
#   4. ncu --set full of the large-frame kernels (seed sort, region growing with sorted seeds) on one 1920x1080 / patch 5 frame.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
BENCH_SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-latency --no-fhd --no-sequence"
$BENCH_SHORT > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    $BENCH_SHORT > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
$BENCH_SHORT > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'cell_stats|region_grow|label|edge_mask|refine|depth' -s 16 -c 6 \
    -f -o $OUT/prof_$TAG $BENCH_SHORT > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
DPX_FUSE_LABELING=0 $BENCH_SHORT > $OUT/plain3_$TAG.log 2>&1 &&
DPX_FUSE_LABELING=0 ncu --set full --clock-control none -k regex:'label' -s 4 -c 2 \
    -f -o $OUT/prof_unfused_$TAG $BENCH_SHORT > $OUT/ncu_unfused_$TAG.log 2>&1
echo "ncu unfused rc=$?"
python tools/frame_profile.py synth 1080 1920 5 > $OUT/plain4_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'seed_|region_grow' -s 12 -c 3 \
    -f -o $OUT/prof_fhd5_$TAG python tools/frame_profile.py synth 1080 1920 5 > $OUT/ncu_fhd5_$TAG.log 2>&1
echo "ncu fhd5 rc=$?"
ls -la $OUT/*$TAG* | tail -12
