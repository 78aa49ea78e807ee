"""Host-side mirror of the reference's model/point_transformer_layer.py.

Same public names, argument meaning and error behaviour (``square_distance``,
``kNN_torch``, ``index_points``, ``PointTransformerLayer``); the arithmetic runs in
libo4d.so.  The reference's ``kNN`` (open3d, lines 33-73) is dead code and is not mirrored.
"""
import torch
from torch import nn

from . import ops


def _wants_grad(module, *tensors):
    """True when the call must be differentiable: grad mode is on and a parameter or an input
    requires grad.  Such calls go through o4d.autograd (forward + backward kernels with saved
    activations); everything else takes the fused inference kernels."""
    if not torch.is_grad_enabled():
        return False
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        return True
    if any(p.requires_grad for p in module.parameters()):
        return True
    # nn.DataParallel replicas (train.py:305) carry their weights as plain attributes of the sub-modules, not as
    # registered parameters: parameters() is empty there although the tensors require grad.
    for m in module.modules():
        for t in m.__dict__.get('_parameters', {}).values():
            if isinstance(t, torch.Tensor) and t.requires_grad:
                return True
        for name in ('weight', 'bias'):
            t = m.__dict__.get(name, None)
            if isinstance(t, torch.Tensor) and t.requires_grad:
                return True
    return False


def square_distance(src, dst):
    """(B,N,3),(B,M,3) -> (B,N,M), reference lines 16-30.  Provided for API completeness only:
    it materialises the full matrix (through kNN with k = M), which the kernels never do."""
    assert src.dim() == 3 and dst.dim() == 3 and src.shape[0] == dst.shape[0]
    out = []
    for b in range(src.shape[0]):
        m = dst.shape[1]
        if m > ops.MAX_K:
            raise NotImplementedError('square_distance: dense (N, M) distances are not part of the '
                                      'hot path; use kNN_torch')
        idx, d = ops.knn(src[b], dst[b], m, sqrt_dist=False, return_dist=True)
        full = torch.empty_like(d)
        full.scatter_(1, idx, d)
        out.append(full)
    return torch.stack(out)


def kNN_torch(query, dataset, k):
    """(B,N0,3),(B,N1,3) -> (B,N0,k) int64 nearest dataset indices, reference lines 76-99.
    Ties resolve to the lower index (the reference's unstable argsort leaves them open)."""
    assert query.dim() == 3 and dataset.dim() == 3, "Input tensors should be 3D."
    assert query.shape[0] == dataset.shape[0], "Input tensors should have same batch size."
    assert query.shape[2] == dataset.shape[2], "Input tensors should have same dimension."
    return torch.stack([ops.knn(query[b], dataset[b], k) for b in range(query.shape[0])])


def index_points(points, idx):
    """Row gather (B,N,C),(B,S,[K]) -> (B,S,[K],C), reference lines 102-113.  Pure data
    movement; inside the kernels the gather is fused and this helper is never called."""
    raw_size = idx.size()
    idx = idx.reshape(raw_size[0], -1)
    res = torch.gather(points, 1, idx[..., None].expand(-1, -1, points.size(-1)))
    return res.reshape(*raw_size, -1)


class PointTransformerLayer(nn.Module):
    """Vector attention over k nearest neighbours (reference lines 116-183)."""

    def __init__(self, dim, pos_mlp_hidden_dim=32, attn_mlp_hidden_mult=2,
                 num_neighbors=16, dim2=None):
        super().__init__()
        if pos_mlp_hidden_dim != 32 or attn_mlp_hidden_mult != 2:
            raise NotImplementedError('o4d kernels are specialised for pos_mlp_hidden_dim=32, '
                                      'attn_mlp_hidden_mult=2 (the only values the reference uses)')
        self.num_neighbors = num_neighbors
        if dim2 is None:
            dim2 = dim
        # creation order = the reference's (RNG stream and state_dict layout depend on it)
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(dim2, dim, bias=False)
        self.to_v = nn.Linear(dim2, dim, bias=False)
        self.pos_mlp = nn.Sequential(
            nn.Linear(3, pos_mlp_hidden_dim), nn.ReLU(), nn.Linear(pos_mlp_hidden_dim, dim))
        self.attn_mlp = nn.Sequential(
            nn.Linear(dim, dim * attn_mlp_hidden_mult), nn.ReLU(),
            nn.Linear(dim * attn_mlp_hidden_mult, dim))
        self.o4d_precision = None  # None = library default (ops.default_precision)

    def o4d_params(self):
        """The 11 tensors in the order o4d_pt_layer_forward expects."""
        return [self.to_q.weight, self.to_k.weight, self.to_v.weight,
                self.pos_mlp[0].weight, self.pos_mlp[0].bias, self.pos_mlp[2].weight, self.pos_mlp[2].bias,
                self.attn_mlp[0].weight, self.attn_mlp[0].bias, self.attn_mlp[2].weight,
                self.attn_mlp[2].bias]

    def forward(self, x, pos, x2=None, pos2=None):
        """x (B,N,D), pos (B,N,3) [x2 (B,M,D2), pos2 (B,M,3)] -> (B,N,D)."""
        if _wants_grad(self, x, x2):
            from . import autograd
            return torch.stack([autograd.pt_layer_train(
                self, x[b], pos[b], None if x2 is None else x2[b], None if pos2 is None else pos2[b])
                for b in range(x.shape[0])])
        params = self.o4d_params()
        out = []
        for b in range(x.shape[0]):
            out.append(ops.pt_layer_forward(
                params, x[b], pos[b], None if x2 is None else x2[b], None if pos2 is None else pos2[b],
                self.num_neighbors, self.o4d_precision))
        return torch.stack(out)
