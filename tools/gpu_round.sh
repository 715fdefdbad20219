#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, then the ncu launch list and full captures of the
# stage kernels (B200_PROFILING.md recipe).  Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01b'
# Everything lands in gpurun_out/; summaries worth keeping are copied into profiles/ afterwards.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
BENCH_SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-latency"

python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
python bench.py > $OUT/bench_$TAG.log 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.log 2>&1; echo "ref rc=$?"; cat $OUT/bench_ref_$TAG.log

if [ "${SKIP_NCU:-0}" = "1" ]; then exit 0; fi
$BENCH_SHORT > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    $BENCH_SHORT > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
$BENCH_SHORT > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'cell_stats|region_grow|label|edge_mask|refine|depth' -s 16 -c 6 \
    -f -o $OUT/prof_$TAG $BENCH_SHORT > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT | tail -20
