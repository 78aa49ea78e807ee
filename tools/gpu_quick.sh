#!/bin/bash
# Short GPU visit: unit test of the fused multi-layer kernel, accuracy table on the released checkpoints, full parity suite,
# short bench lines (A/B switches through the environment), synccheck on the tcgen05 kernels.
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_residual_block" > $OUT/unit.log 2>&1; echo "unit exit $?"; tail -4 $OUT/unit.log
timeout 300 python tools/ckpt_error.py > $OUT/ckpt_error.json 2>&1; echo "ckpt exit $?"
python - <<PY
import json
try:
    d=json.load(open('$OUT/ckpt_error.json'))
    for w in d:
        print(w, 'ref', '%.2e'%d[w]['ref_fp32_vs_fp64'], {k:('%.2e'%v['out_vs_fp64'],'%.2e'%v['penult_vs_fp64'],'%.2f ms'%v['ms_65536_queries']) for k,v in d[w].items() if isinstance(v,dict)})
except Exception as e:
    print('ckpt table unreadable', e); print(open('$OUT/ckpt_error.json').read()[-1500:])
PY
timeout 900 python -m pytest tests -m gpu -x -q -rs -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "checkpoint:|passed|failed|skipped|Error|error|assert" $OUT/pytest_gpu.log | tail -12
for variant in "default" $AB; do
  envs=""; [ "$variant" != "default" ] && envs="$variant"
  env $envs timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-train-step --no-strong-scaling --no-carla > $OUT/bench_$variant.json 2> $OUT/bench_$variant.err; echo "bench [$variant] exit $?"
  python -c "
import json,sys
l=json.loads(open('$OUT/bench_$variant.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print({k:(round(v['ms'],2),v['launches'],round(v['tflops'])) for k,v in l['kernel_families'].items()}); print('roofline frac', l['roofline']['frac'])"
done
if [ "$SKIP_SAN" != "1" ]; then
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -x -q -k "fused_residual_block and 1024 or fused_attention_decoder_shapes" > $OUT/synccheck.log 2>&1; echo "synccheck exit $?"
grep -E "ERROR SUMMARY|passed|failed|at void" $OUT/synccheck.log | sort | uniq -c | tail -8
fi
if [ -f occlusions-4d_b200/o4d/libo4d_stamps.so ] && [ "$SKIP_STAMPS" != "1" ]; then
O4D_LIB=$PWD/occlusions-4d_b200/o4d/libo4d_stamps.so timeout 300 python tools/stamps_fused.py > $OUT/stamps.log 2>&1; grep -v linear $OUT/stamps.log
fi
