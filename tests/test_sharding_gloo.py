"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous query sharding plus the
all-gather of output shards must reproduce the single-rank result bit for bit.  The
per-shard decode here is the CPU oracle (test infrastructure) -- on the GPU box the same
plumbing runs with the CUDA decoder and NCCL (tests/test_gpu_models.py, bench.py)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    for p in (REPO, os.path.join(REPO, 'occlusions-4d_b200')):
        if p not in sys.path:
            sys.path.insert(0, p)
    from o4d import parallel
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(3)
    q = torch.rand(1001, 4)
    w = torch.rand(4, 5)

    def decode(batch):  # any row-wise function: sharding must not change its results
        return torch.sin(batch @ w)

    full = parallel.decode_sharded(decode, q, batch_size=128)
    single = torch.cat([decode(q[s:s + 128]) for s in range(0, q.shape[0], 128)])
    a, b = parallel.shard_range(q.shape[0], rank, world)
    ok = torch.equal(full, single) and full.shape == (1001, 5)
    # ragged + empty shards
    tiny = parallel.decode_sharded(decode, q[:1], batch_size=128)
    ok = ok and torch.equal(tiny, decode(q[:1]))
    # data-parallel gradient averaging (training, config 5): every rank ends with the mean over
    # ranks, parameters without a local gradient count as zeros, buckets cover every parameter
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(7, 9), torch.nn.Linear(9, 3), torch.nn.Linear(3, 2))
    local = []
    for i, p in enumerate(net.parameters()):
        if i == 3 and rank == 1:
            local.append(None)
            continue
        p.grad = torch.full_like(p, float(rank + 1)) * (i + 1)
        local.append(p.grad.clone())
    nb = parallel.allreduce_gradients([net], bucket_bytes=256)
    ok = ok and nb >= 2
    for i, p in enumerate(net.parameters()):
        want = (1 + 2) * (i + 1) / 2.0 if i != 3 else 1 * (i + 1) / 2.0
        ok = ok and bool(torch.allclose(p.grad, torch.full_like(p, want)))
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_decode_equals_single_rank_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
