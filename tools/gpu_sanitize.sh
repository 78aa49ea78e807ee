#!/bin/bash
# compute-sanitizer pass over the small-shape GPU tests (memcheck, then racecheck + synccheck on the kernels that
# use shared memory / cluster barriers).  Slow (10-100x): run on the kernel tests only, never on the full-size ones.
# Usage (from the repo root, under gpurun): bash tools/gpu_sanitize.sh <tag> [pytest -k expression]
TAG=${1:-san}
SEL=${2:-"knn or fps or filter or bounds or loss or grid or activation or local_blend or posenc"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export O4D_PRECISION=${O4D_PRECISION:-1}
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_kernels.py tests/test_gpu_sampler.py -m gpu -x -q -k "$SEL and not training_size" \
      > $OUT/$tool.log 2>&1
  echo "$tool exit $?" | tee -a $OUT/$tool.log
  grep -E "ERROR SUMMARY|passed|failed" $OUT/$tool.log | tail -3
done
