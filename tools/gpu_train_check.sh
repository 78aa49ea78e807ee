#!/bin/bash
# GPU visit for the training path: backward parity tests (all failures collected, no -x).
TAG=${1:-train}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_train.py -q -s -m gpu > $OUT/pytest_train.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_train.log
tail -60 $OUT/pytest_train.log
