#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the bench under torchrun (weak-scaling headline, strong-scaling frames with the NCCL
# bit-exactness check, second-device-in-one-process check), then the CPU reference arm launched the same way.
N=${1:-2}
TAG=${2:-multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "bench exit $?"
python - <<PY
import json
l=json.loads(open('$OUT/bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','n_gpus','ms_per_step','gpu_launches','clocks')}); print('e2e', l['e2e']); print('strong', json.dumps(l['strong_scaling'])[:1200]); print('second device', l['second_device_in_process']); print('train', {k:v for k,v in l['train_step'].items() if k in ('ms_per_step','query_grads_per_s','samples_per_step')})
PY
tail -3 $OUT/bench_${N}gpu.err
