"""TEST INFRASTRUCTURE ONLY -- oracle for the two third-party ops on the hot path.

The reference calls ``torch_cluster.fps`` (model/modules.py:133-134) and
``torch_cluster.knn`` (model/modules.py:142-146).  ``torch_cluster``
(rusty1s/pytorch_cluster, no version pinned by the reference: README.md:32 only says
"libraries imported in __init__.py") is NOT vendored under /root/reference and is not
installed in this image, so its results cannot be pinned here:

    *** parity unpinned at this boundary ***

What follows restates the *published* algorithm of those two ops as the reference's call
sites use them:

fps(src (B*N,3), batch, ratio, random_start)
    per batch segment: n_out = ceil(ratio * n) samples; first sample = first point of
    the segment (random_start=False) or a uniformly random one (True); every further
    sample = argmax over the running min squared distance to the already selected
    set (first maximum on ties); indices are global (flattened) and in selection
    order -- the caller sorts them (modules.py:135).
knn(x, y, k, batch_x, batch_y)
    for every y the k nearest x of the same batch id by Euclidean distance; returns
    (2, |y|*k) with row 0 = y index, row 1 = x index.  The caller only uses row 1
    and max-pools over it (modules.py:149-158) so the order within a row is free.
"""
import math

import numpy as np
import torch


def _sqdist_rows(p, c):
    """d2[i] = ((dx*dx + dy*dy) + dz*dz) in fp32, the canonical order of this repo."""
    d = p - c
    d = d * d
    return (d[:, 0] + d[:, 1]) + d[:, 2]


def fps_segment(xyz, n_out, start=0):
    """Farthest point sampling of one cloud. xyz (N,3) fp32 -> (n_out,) int64, selection order."""
    xyz = xyz.detach().to(torch.float32).cpu()
    n = xyz.shape[0]
    sel = torch.empty(n_out, dtype=torch.int64)
    mind = torch.full((n,), float('inf'), dtype=torch.float32)
    cur = int(start)
    for i in range(n_out):
        sel[i] = cur
        d2 = _sqdist_rows(xyz, xyz[cur][None, :])
        mind = torch.minimum(mind, d2)
        cur = int(torch.argmax(mind))  # first maximum on ties (torch.argmax on CPU).
    return sel


def fps(src, batch=None, ratio=0.5, random_start=True):
    """Call-compatible with torch_cluster.fps as used at modules.py:133-134."""
    dev = src.device
    src_c = src.detach().cpu()
    if batch is None:
        batch = torch.zeros(src_c.shape[0], dtype=torch.int64)
    batch = batch.detach().cpu()
    out = []
    for b in torch.unique(batch).tolist():
        ids = torch.nonzero(batch == b).flatten()
        n = ids.numel()
        # upstream multiplies the segment size by ratio IN src.dtype, then ceils
        # (228 * fp32(1/3) rounds to exactly 76.0; in fp64 it would ceil to 77).
        n_out = int(torch.ceil(torch.tensor(n, dtype=src_c.dtype) *
                               torch.tensor(ratio, dtype=src_c.dtype)))
        start = int(torch.randint(0, n, (1,))) if random_start else 0
        sel = fps_segment(src_c[ids], n_out, start)
        out.append(ids[sel])
    return torch.cat(out).to(dev)


def knn_bruteforce(query_xyz, ref_xyz, k, sqrt=False):
    """Canonical kNN of this repo: ascending (distance, index); fp32 ((dx2+dy2)+dz2).

    query (N,3), ref (M,3) -> idx (N,k) int64, dist (N,k) fp32 (squared unless sqrt).
    """
    q = query_xyz.detach().to(torch.float32).cpu().numpy()
    r = ref_xyz.detach().to(torch.float32).cpu().numpy()
    n = q.shape[0]
    idx = np.empty((n, k), dtype=np.int64)
    dist = np.empty((n, k), dtype=np.float32)
    step = max(1, (1 << 24) // max(1, r.shape[0]))
    for s in range(0, n, step):
        d = q[s:s + step, None, :] - r[None, :, :]
        d = d * d
        d2 = (d[..., 0] + d[..., 1]) + d[..., 2]
        if sqrt:
            d2 = np.sqrt(d2)
        order = np.argsort(d2, axis=1, kind='stable')[:, :k]
        idx[s:s + step] = order
        dist[s:s + step] = np.take_along_axis(d2, order, axis=1)
    return torch.from_numpy(idx), torch.from_numpy(dist)


def knn(x, y, k, batch_x=None, batch_y=None):
    """Call-compatible with torch_cluster.knn as used at modules.py:142-146."""
    dev = x.device
    xc, yc = x.detach().cpu(), y.detach().cpu()
    if batch_x is None:
        batch_x = torch.zeros(xc.shape[0], dtype=torch.int64)
    if batch_y is None:
        batch_y = torch.zeros(yc.shape[0], dtype=torch.int64)
    batch_x, batch_y = batch_x.cpu(), batch_y.cpu()
    rows, cols = [], []
    for b in torch.unique(batch_y).tolist():
        ix = torch.nonzero(batch_x == b).flatten()
        iy = torch.nonzero(batch_y == b).flatten()
        nbr, _ = knn_bruteforce(yc[iy, :3], xc[ix, :3], k)
        rows.append(iy.repeat_interleave(k))
        cols.append(ix[nbr.reshape(-1)])
    return torch.stack([torch.cat(rows), torch.cat(cols)]).to(dev)
