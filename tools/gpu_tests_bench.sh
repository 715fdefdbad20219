#!/bin/bash
# GPU tests + smoke + both bench arms, no profiler.  gpurun --timeout 1800 -- 'bash tools/gpu_tests_bench.sh r02a'
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L; nproc; free -g | head -2
python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu_$TAG.log
python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
python bench.py ${BENCH_ARGS:---steps 20 --warmup 5} > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -5 $OUT/bench_$TAG.err; head -c 3000 $OUT/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>&1; echo "ref rc=$?"; head -c 600 $OUT/bench_ref_$TAG.json
