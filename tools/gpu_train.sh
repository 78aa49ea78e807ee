#!/bin/bash
# GPU visit for the training path: backward parity tests, one config-5-shape decoder frame timed (bf16x3 and single pass),
# ncu launch list of two passes summarised per kernel.  Usage: bash tools/gpu_train.sh <tag>; env SKIP_TESTS=1, SKIP_NCU=1
TAG=${1:-train}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "$SKIP_TESTS" != "1" ]; then
timeout 1200 python -m pytest tests/test_gpu_train.py -q -m gpu > $OUT/pytest_train.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_train.log
fi
timeout 300 python tools/prof_train.py 6 1 > $OUT/time_p1.log 2>&1; tail -1 $OUT/time_p1.log
timeout 300 python tools/prof_train.py 6 2 > $OUT/time_p2.log 2>&1; tail -1 $OUT/time_p2.log
if [ "$SKIP_NCU" != "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_train.csv python tools/prof_train.py 2 ${NCU_PREC:-1} > $OUT/ncu_train.log 2>&1; echo "ncu exit $?"
python tools/ncu_summary.py $OUT/launches_train.csv "train frame" > $OUT/launches_train.md; head -24 $OUT/launches_train.md
fi
