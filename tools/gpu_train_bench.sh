#!/bin/bash
# Training path: parity tests + the config-5 step of bench.py (train_step key only).  Usage: bash tools/gpu_train_bench.sh <tag>
TAG=${1:-trainb}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_kernels.py -q -m gpu -x > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-strong-scaling --no-carla > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python -c "
import json
l=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1])
t=l['train_step']; print('train step ms', t['ms_per_step'], 'bf16', t['bf16']['ms_per_step'], 'loss', t['loss'], t['bf16']['loss'])
print('inference', l['value'], l['ms_per_step'], {k:round(v['ms'],2) for k,v in l['kernel_families'].items()})"
timeout 300 python tools/prof_train_step.py 1 2>&1 | cut -c1-72,142-200 | head -30
