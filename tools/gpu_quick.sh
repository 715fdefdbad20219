#!/bin/bash
# Quick GPU pass: parity tests + bench at several stage-1 warp counts (A/B), no profiler.
set -u
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu_$TAG.log
for w in ${WARPS_LIST:-12 16 8}; do
  DPX_STREAM_WARPS=$w python bench.py --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} > $OUT/bench_${TAG}_w$w.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${TAG}_w$w.log").read().strip().splitlines()[-1])
    s = d["roofline"]["stages"]
    print("warps=$w value=%.0f fps ms/step=%.4f" % (d["value"], d["ms_per_step"]), {k: round(v["ms"], 4) for k, v in s.items()}, "cell_stats frac=%.3f" % s["cell_stats"]["frac"])
except Exception as e:
    print("warps=$w failed", e); print(open("$OUT/bench_${TAG}_w$w.log").read()[-2000:])
PY
done
