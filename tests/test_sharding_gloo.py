"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous query sharding plus the
all-gather of output shards must reproduce the single-rank result bit for bit.  The
per-shard decode here is the CPU oracle (test infrastructure) -- on the GPU box the same
plumbing runs with the CUDA decoder and NCCL (tests/test_gpu_models.py, bench.py)."""
import os
import sys

import pytest
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


def _warmup_worker(rank, world, port, ret):
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    import bench
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    # every step holds a collective (the gradient all-reduce of the training step); the two ranks' allocators settle at
    # different steps: rank 0 after 3, rank 1 after 6
    quiet_after = 3 if rank == 0 else 6
    state = {'steps': 0, 'allocs': 0}

    def step():
        state['steps'] += 1
        state['allocs'] += 20 if state['steps'] <= quiet_after else 0
        t = torch.ones(4)
        dist.all_reduce(t)                       # mismatched step counts would pair this with the barrier below

    n = bench.warm_up_until_quiet(step, lambda: state['allocs'], world, 'cpu')
    dist.barrier()
    ret[rank] = (n, state['steps'])
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_adaptive_warm_up_runs_the_same_number_of_steps_on_every_rank_gloo():
    """bench.train_step_bench warms up until the caching allocator is quiet; the step count must be agreed across
    ranks, or the ranks' collectives pair up wrongly and hang (it did, once, on 2 GPUs)."""
    world = 2
    port = 31500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_warmup_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: (7, 7), 1: (7, 7)}      # rank 1 needs 6 noisy steps + 1 quiet one; rank 0 follows
    # single process: stops on its own count, bounded by max_steps
    import bench
    calls = {'n': 0, 'a': 0}

    def step():
        calls['n'] += 1
        calls['a'] += 5
    assert bench.warm_up_until_quiet(step, lambda: calls['a'], 1, 'cpu') == 12
