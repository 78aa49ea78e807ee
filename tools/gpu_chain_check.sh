#!/bin/bash
# First GPU visit of the fused multi-layer kernel: unit test, model parity, A/B bench, sanitizer on the small cases.
TAG=${1:-chain}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_residual_block" > $OUT/unit.log 2>&1; echo "unit exit $?"; tail -15 $OUT/unit.log
timeout 900 python -m pytest tests -m gpu -x -q -rs -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "checkpoint:|passed|failed|skipped|Error|error|assert" $OUT/pytest_gpu.log | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-train-step --no-strong-scaling > $OUT/bench_chain.json 2> $OUT/bench_chain.err; echo "bench exit $?"
python -c "
import json,sys
l=json.loads(open('$OUT/bench_chain.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(l['kernel_families']); print(l['e2e']); print(l['carla_config3'])"
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -x -q -k "fused_residual_block and 1024 or fused_attention_decoder_shapes" > $OUT/synccheck.log 2>&1; echo "synccheck exit $?"
grep -E "ERROR SUMMARY|passed|failed|at void" $OUT/synccheck.log | sort | uniq -c | tail -8
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_residual_block and (1024 or 3001)" > $OUT/memcheck.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed|at void" $OUT/memcheck.log | sort | uniq -c | tail -8
