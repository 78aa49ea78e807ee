"""Host-side mirror of the reference's model/modules.py (PointTransformerBlock,
DownTransition).  ``UpTransition`` (reference lines 166-289) is dead in both released
configurations (enable_decoder=False, train.py:223) and is not part of the hot path.
"""
import math

import torch
from torch import nn

from . import ops
from . import point_transformer_layer
from .point_transformer_layer import _wants_grad


class PointTransformerBlock(nn.Module):
    """Linear + point transformer layer + Linear, residual (reference lines 18-67)."""

    def __init__(self, d_in, d_hidden, d_out, num_neighbors=16, d_hidden_abstract=None):
        super().__init__()
        self.d_in = d_in
        self.d_hidden = d_hidden
        self.d_out = d_out
        self.num_neighbors = num_neighbors
        self.layer1 = nn.Linear(d_in, d_hidden)
        self.layer2 = point_transformer_layer.PointTransformerLayer(
            d_hidden, pos_mlp_hidden_dim=32, attn_mlp_hidden_mult=2,
            num_neighbors=num_neighbors, dim2=d_hidden_abstract)
        self.layer3 = nn.Linear(d_hidden, d_out)
        self.o4d_precision = None

    def o4d_params(self):
        """The 15 tensors in the order o4d_pt_block_forward expects (= state_dict order)."""
        return [self.layer1.weight, self.layer1.bias] + self.layer2.o4d_params() + \
               [self.layer3.weight, self.layer3.bias]

    def forward(self, x, p, x2=None, p2=None):
        """x (B,N,d_in), p (B,N,3) [x2 (B,M,d2), p2 (B,M,3)] -> (z (B,N,d_out), p)."""
        assert x.shape[:2] == p.shape[:2]
        if x2 is not None:
            assert x2.shape[:2] == p2.shape[:2]
        if not (self.d_in == self.d_hidden == self.d_out):
            raise NotImplementedError('o4d: PointTransformerBlock needs d_in == d_hidden == d_out '
                                      '(true for every block the reference builds)')
        if _wants_grad(self, x, x2):
            from . import autograd
            z = [autograd.pt_block_train(self, x[b], p[b], None if x2 is None else x2[b],
                                         None if p2 is None else p2[b]) for b in range(x.shape[0])]
            return (torch.stack(z), p)
        params = self.o4d_params()
        z = []
        for b in range(x.shape[0]):
            z.append(ops.pt_block_forward(
                params, x[b], p[b], None if x2 is None else x2[b], None if p2 is None else p2[b],
                self.num_neighbors, self.o4d_precision))
        return (torch.stack(z), p)


class DownTransition(nn.Module):
    """Farthest point sampling + kNN / MLP + local max pooling (reference lines 70-163)."""

    def __init__(self, d_in, d_out, factor=2, knn_k=8, norm_type='none', fps_random_start=True):
        super().__init__()
        self.d_in = d_in
        self.d_out = d_out
        self.factor = factor
        self.knn_k = knn_k
        self.norm_type = norm_type
        self.fps_random_start = fps_random_start
        if norm_type == 'none':
            self.mlp = nn.Sequential(nn.Linear(d_in, d_out), nn.ReLU())
        elif norm_type == 'batch':
            self.mlp = nn.Sequential(nn.Linear(d_in, d_out), nn.BatchNorm1d(d_out, eps=1e-3), nn.ReLU())
        elif norm_type == 'layer':
            self.mlp = nn.Sequential(nn.Linear(d_in, d_out), nn.LayerNorm(d_out), nn.ReLU())
        else:
            raise ValueError()
        self.o4d_precision = None

    def o4d_params(self):
        if self.norm_type == 'layer':
            return [self.mlp[0].weight, self.mlp[0].bias, self.mlp[1].weight, self.mlp[1].bias]
        return [self.mlp[0].weight, self.mlp[0].bias, None, None]

    def o4d_start(self, n):
        """FPS start index: 0 at test time (inference.py:59), uniform random in training."""
        return int(torch.randint(0, n, (1,))) if self.fps_random_start else 0

    def forward(self, x, p):
        """x (B,N,d_in), p (B,N,3) -> (z (B,ceil(N/factor),d_out), p_sub (B,ceil(N/factor),3))."""
        assert x.shape[:2] == p.shape[:2]
        if self.norm_type == 'batch':
            raise NotImplementedError("o4d: norm_type 'batch' is unused by the released configurations")
        (B, N, _) = x.shape
        assert int(math.ceil(N / self.factor)) >= 1
        norm = 1 if self.norm_type == 'layer' else 0
        zs, ps = [], []
        if _wants_grad(self, x):
            from . import autograd
            for b in range(B):
                z, p_sub = autograd.down_train(self, x[b], p[b].contiguous(), self.o4d_start(N))
                zs.append(z)
                ps.append(p_sub)
            return (torch.stack(zs), torch.stack(ps))
        for b in range(B):
            z, p_sub = ops.down_forward(self.o4d_params(), x[b], p[b], self.d_out, self.factor, self.knn_k,
                                        norm, self.o4d_start(N), self.o4d_precision)
            zs.append(z)
            ps.append(p_sub)
        return (torch.stack(zs), torch.stack(ps))
