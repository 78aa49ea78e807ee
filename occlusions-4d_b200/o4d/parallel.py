"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the
CPU tests).  Queries are independent given the scene encoding (implicit.py:413-443 has no
cross-query operation), so a frame's queries shard contiguously across ranks with no
data-path collective; the only exchange is the all-gather of the (N/G, d_out) output shards.
The reference has no multi-GPU inference at all (eval/test.py:156-157 pins one device).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced split of n units: sizes differ by at most one."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_rows(local, n_total, group=None):
    """All-gather row shards produced with shard_range into one (n_total, C) tensor on every rank.
    Shards are padded to the largest size so a single fixed-size all_gather is used."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    width = local.shape[1]
    max_rows = -(-n_total // world)
    padded = local
    if local.shape[0] < max_rows:
        padded = torch.zeros((max_rows, width), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    gathered = torch.empty((world * max_rows, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(n_total, r, world)
        parts.append(gathered[r * max_rows:r * max_rows + (b - a)])
    return torch.cat(parts, dim=0)


def decode_sharded(decode_fn, query_all, batch_size, group=None):
    """Decode this rank's contiguous shard of `query_all` in mini-batches with
    decode_fn(query_batch) -> (n, C) and return the assembled (N, C) result on every rank."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    a, b = shard_range(query_all.shape[0], rank, world)
    outs = [decode_fn(query_all[s:min(s + batch_size, b)]) for s in range(a, b, batch_size)]
    if outs:
        local = torch.cat(outs, dim=0)
    else:
        probe = decode_fn(query_all[:0])
        local = probe.reshape(0, probe.shape[-1])
    if world == 1:
        return local
    return all_gather_rows(local, query_all.shape[0], group)


def allreduce_gradients(modules, group=None, bucket_bytes=32 << 20):
    """Data-parallel gradient averaging, one sample per GPU (replaces nn.DataParallel at
    train.py:305: the reference's loss is the mean over replicas of per-replica means,
    pipeline.py:149-153 / loss.py:272-275, so gradients are AVERAGED over ranks).
    Gradients are packed into flat buckets of about `bucket_bytes` (7.26 M fp32 parameters =
    one 29 MB bucket for the released models), each reduced with one all_reduce(SUM) over
    NCCL / NVLink and scaled by 1/world.  Parameters without a gradient on this rank
    contribute zeros (every rank must walk the same parameter list).  Returns bucket count."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    params = [p for m in modules for p in m.parameters() if p.requires_grad]
    buckets, cur, cur_bytes = [], [], 0
    for p in params:
        nbytes = p.numel() * p.element_size()
        if cur and cur_bytes + nbytes > bucket_bytes:
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(p)
        cur_bytes += nbytes
    if cur:
        buckets.append(cur)
    for bucket in buckets:
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for p in bucket:
            n = p.numel()
            if p.grad is None:
                p.grad = flat[off:off + n].reshape(p.shape).clone()
            else:
                p.grad.copy_(flat[off:off + n].reshape(p.shape))
            off += n
    return len(buckets)
