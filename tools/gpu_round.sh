#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list, ncu full capture of the two top kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
tail -c 3000 $OUT/bench.json
if [ "$SKIP_REF" != "1" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"
cat $OUT/bench_reference.json
fi
if [ "$SKIP_NCU" != "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train-step > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_fused_kernel|linear_tc_kernel' -s 40 -c 4 \
    -o $OUT/prof_top python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train-step > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $OUT
fi
