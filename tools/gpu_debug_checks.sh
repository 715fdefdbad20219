#!/bin/bash
# The hot-path cases of tools/sanitizer_cases.py and the parity tests against the DEBUG build of the library
# (make -C deplex_b200/csrc debug: the kernels' own bounds / invariant checks, common.cuh DPX_CHECK).  compute-sanitizer
# is closed on the B200 pool this repository is developed on (profiles/r02_sanitizer_refused.txt); this is its stand-in.
#   gpurun --timeout 1500 -- 'bash tools/gpu_debug_checks.sh r02'
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export DPX_LIB_PATH=$PWD/deplex_b200/libdeplex_b200_dbg.so
ls -la $DPX_LIB_PATH || exit 1
python tools/sanitizer_cases.py > $OUT/debug_checks_cases_$TAG.log 2>&1; echo "cases rc=$?"; cat $OUT/debug_checks_cases_$TAG.log
python -m pytest tests/test_parity_gpu.py tests/test_pipeline_gpu.py tests/test_refine_gpu.py -m gpu -q -x > $OUT/debug_checks_pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/debug_checks_pytest_$TAG.log
grep -c "DPX_CHECK failed" $OUT/debug_checks_cases_$TAG.log $OUT/debug_checks_pytest_$TAG.log
python - <<'PY'
import os
from deplex_b200 import _capi
_capi.load()
print("library under test:", [l.split()[-1] for l in open("/proc/self/maps") if "libdeplex_b200" in l][0])
PY
