#!/bin/bash
# compute-sanitizer pass: memcheck over the small-shape kernel tests, then racecheck + synccheck over tests that drive the
# kernels with hand-rolled mbarrier / named-barrier protocols (tcgen05 dense layer, fused attention, cluster FPS) at small
# shapes.  Slow (10-100x): never on the full-size tests.
# Usage (from the repo root, under gpurun): bash tools/gpu_sanitize.sh <tag>
TAG=${1:-san}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export O4D_PRECISION=${O4D_PRECISION:-1}
SEL_SMALL="knn or fps or filter or bounds or loss or grid or activation or local_blend or posenc"
SEL_TC="(linear and not backward) or fused_attention_decoder_shapes or fps"
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_sampler.py -m gpu -x -q -k "($SEL_SMALL) and not training_size" > $OUT/memcheck.log 2>&1
echo "memcheck exit $?" | tee -a $OUT/memcheck.log
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -x -q -k "$SEL_TC" > $OUT/memcheck_tc.log 2>&1
echo "memcheck_tc exit $?" | tee -a $OUT/memcheck_tc.log
for tool in racecheck synccheck; do
  timeout 240 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -x -q -k "$SEL_TC" > $OUT/$tool.log 2>&1
  echo "$tool exit $?" | tee -a $OUT/$tool.log
done
# training path: the tcgen05 weight gradient (tensor-map loads, mbarrier rings, named barrier) and the masked dX epilogue
for tool in memcheck synccheck; do
  timeout 240 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "linear_backward and (4100 or 4500 or 4096)" > $OUT/${tool}_wgrad.log 2>&1
  echo "${tool}_wgrad exit $?" | tee -a $OUT/${tool}_wgrad.log
done
grep -E "ERROR SUMMARY|passed|failed| exit " $OUT/*.log | tail -24
