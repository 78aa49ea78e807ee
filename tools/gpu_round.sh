#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list, ncu full capture of the top kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
# Env: SKIP_REF=1 (no reference arm), SKIP_NCU=1, SKIP_SAN=1 (no compute-sanitizer pass), SKIP_TESTS=1, EXTRA="cmd" (run first)
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ -n "$EXTRA" ]; then bash -c "$EXTRA" > $OUT/extra.log 2>&1; echo "extra exit $?"; tail -40 $OUT/extra.log; fi
if [ "$SKIP_TESTS" != "1" ]; then
timeout 1200 python -m pytest tests -m gpu -x -q -rs -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -E "checkpoint:|passed|failed|skipped|Error|error" $OUT/pytest_gpu.log | tail -15
fi
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
tail -c 6000 $OUT/bench.json
if [ "$SKIP_REF" != "1" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"
cat $OUT/bench_reference.json | cut -c1-400
fi
if [ "$SKIP_NCU" != "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train-step --no-torch-gpu-baseline --no-strong-scaling --no-carla > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-attn_fused_kernel|mlp_chain_kernel}" -s 40 -c 4 \
    -o $OUT/prof_top python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train-step --no-torch-gpu-baseline --no-strong-scaling --no-carla > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
fi
if [ "$SKIP_SAN" != "1" ]; then bash tools/gpu_sanitize.sh $TAG/san; fi
ls -la $OUT
