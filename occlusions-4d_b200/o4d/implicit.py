"""Host-side mirror of the reference's model/implicit.py: positional_encode, ResnetBlockFC,
ResnetFC, LocalPclResnetFC.  Same constructor kwargs (stored in checkpoints), parameter
names, creation order and return tuples.  The attention decoder's forward is ONE call
into libo4d.so per query mini-batch (o4d_decoder_forward) plus one per scene
(o4d_decoder_prepare_scene, cached on the abstract cloud's identity).
"""
import math

import torch
from torch import nn

from . import modules
from . import ops
from .point_transformer_layer import _wants_grad


def positional_encode(points, base_frequency, num_powers):
    """(...,4) -> (...,4*(2F+1)) Fourier features, reference lines 20-43 (runs in libo4d.so)."""
    if base_frequency != 0.1:
        raise NotImplementedError('o4d: base_frequency is 0.1 everywhere in the reference '
                                  '(implicit.py:184,405) and is baked into the kernel')
    import ctypes
    from . import _lib
    pts = ops._f32(points, 'points')
    flat = pts.reshape(-1, pts.shape[-1]).contiguous()
    width = flat.shape[1] * (2 * num_powers + 1)
    out = torch.empty((flat.shape[0], width), dtype=torch.float32, device=flat.device)
    rc = _lib.lib().o4d_posenc_f32(ops._ptr(flat), flat.shape[0], flat.shape[1], int(num_powers),
                                   ops._ptr(out), ops._stream(flat))
    _lib.check(rc, 'o4d_posenc_f32')
    return out.reshape(*pts.shape[:-1], width)


class Swish(nn.Module):
    """x * sigmoid(x) (reference lines 46-56).  No kernel: the released configs use ReLU."""

    def forward(self, input):
        raise NotImplementedError("o4d: activation 'swish' has no kernel (released configs use relu)")


def instantiate_activation_fn(activation_str):
    if activation_str == 'relu':
        return nn.ReLU()
    elif activation_str == 'swish':
        return Swish()
    raise ValueError('Unknown activation: ' + str(activation_str))


class ResnetBlockFC(nn.Module):
    """x + fc_1(act(fc_0(act(x)))) (reference lines 68-101)."""

    def __init__(self, d_in=64, d_hidden=256, d_out=64, activation='relu'):
        super().__init__()
        self.d_in = d_in
        self.d_hidden = d_hidden
        self.d_out = d_out
        self.fc_0 = nn.Linear(d_in, d_hidden, bias=True)
        self.fc_1 = nn.Linear(d_hidden, d_out, bias=True)
        self.activation = instantiate_activation_fn(activation)
        self.shortcut = None if d_in == d_out else nn.Linear(d_in, d_out, bias=False)

    def forward(self, x):
        if isinstance(self.activation, Swish):
            self.activation(x)
        if _wants_grad(self, x):
            from . import autograd
            net = autograd.linear(x, self.fc_0.weight, self.fc_0.bias, relu_in=True)
            x_s = x if self.shortcut is None else autograd.linear(x, self.shortcut.weight)
            return autograd.linear(net, self.fc_1.weight, self.fc_1.bias, residual=x_s, relu_in=True)
        if self.shortcut is None:       # the whole block as one launch of the fused multi-layer kernel when it fits
            fused = ops.resblock(x, self.fc_0.weight, self.fc_0.bias, self.fc_1.weight, self.fc_1.bias)
            if fused is not None:
                return fused
        net = ops.linear(x, self.fc_0.weight, self.fc_0.bias, relu_in=True)
        x_s = x if self.shortcut is None else ops.linear(x, self.shortcut.weight)
        return ops.linear(net, self.fc_1.weight, self.fc_1.bias, residual=x_s, relu_in=True)


class ResnetFC(nn.Module):
    """Global-embedding conditioned residual MLP (reference lines 104-208)."""

    def __init__(self, mixed_precision=False, d_in=4, d_hidden=256, d_out=64, d_latent=256,
                 n_blocks=5, pos_encoding_freqs=0, activation='relu'):
        super().__init__()
        self.mixed_precision = mixed_precision
        self.d_in = d_in
        self.d_hidden = d_hidden
        self.d_out = d_out
        self.d_latent = d_latent
        self.n_blocks = n_blocks
        self.pos_encoding_freqs = pos_encoding_freqs
        self.actual_d_in = d_in * (pos_encoding_freqs * 2 + 1) if pos_encoding_freqs > 0 else d_in
        if self.actual_d_in > 0:
            self.lin_in = nn.Linear(self.actual_d_in, d_hidden, bias=True)
        self.lin_out = nn.Linear(d_hidden, d_out, bias=True)
        self.blocks = nn.ModuleList(
            [ResnetBlockFC(d_in=d_hidden, d_hidden=d_hidden, d_out=d_hidden, activation=activation)
             for _ in range(n_blocks)])
        if d_latent > 0:
            self.lin_z = nn.ModuleList(
                [nn.Linear(d_latent, d_hidden, bias=True) for _ in range(n_blocks)])
        self.activation = instantiate_activation_fn(activation)
        self.o4d_precision = None

    def forward(self, points, features):
        return self.do_forward(points, features)

    def do_forward(self, points, features):
        """points (B,N,4), features (B,D) | (B,N,D) -> (output (B,N,G), penult (B,N,H)).
        Secondary mode of the reference (global-only / 'feature' conditioning): composed from
        o4d_linear_f32 calls, one per layer."""
        if isinstance(self.activation, Swish):
            self.activation(points)
        lin = ops.linear
        if _wants_grad(self, points, features):
            from . import autograd
            lin = autograd.linear
        if len(points.shape) == 2:
            points = points.unsqueeze(0)
            features = features.unsqueeze(0)
            no_batch = True
        else:
            no_batch = False
        assert points.shape[0] == features.shape[0]
        (B, N, _) = points.shape
        if len(features.shape) != 2:
            assert points.shape[1] == features.shape[1]
        assert points.shape[-1] == self.d_in
        assert features.shape[-1] == self.d_latent
        if self.d_in <= 0:
            raise NotImplementedError('o4d: d_in == 0 is not used by the reference callers')
        if self.pos_encoding_freqs > 0:
            points = positional_encode(points, 0.1, self.pos_encoding_freqs)
        x = lin(points, self.lin_in.weight, self.lin_in.bias)
        for blkid in range(self.n_blocks):
            if self.d_latent > 0:
                if len(features.shape) == 2:
                    z = lin(features, self.lin_z[blkid].weight, self.lin_z[blkid].bias)
                    x = x + z.unsqueeze(1).expand_as(x)
                else:   # per-point features: the add rides on the layer's residual input
                    x = lin(features, self.lin_z[blkid].weight, self.lin_z[blkid].bias, residual=x)
            x = self.blocks[blkid](x)
        penult = x
        output = lin(x, self.lin_out.weight, self.lin_out.bias, relu_in=True)
        if no_batch:
            output = output.squeeze(0)
            penult = penult.squeeze(0)
        return (output, penult)


class LocalPclResnetFC(ResnetFC):
    """ResnetFC + local feature conditioning + query-to-abstract cross attention
    (reference lines 211-445)."""

    def __init__(self, num_local_features=0, local_mode='attention', d_latent_local=64,
                 cross_attn_neighbors=12, cross_attn_layers=1, cr_attn_type='cccccccccc', **kwargs):
        super().__init__(**kwargs)
        self.num_local_features = num_local_features
        self.local_mode = local_mode
        self.d_latent_local = d_latent_local
        self.cross_attn_neighbors = cross_attn_neighbors
        self.cross_attn_layers = cross_attn_layers
        self.cr_attn_type = cr_attn_type
        if local_mode == 'attention':
            pt_blocks = []
            use_pt_inds = []
            for pt_idx in range(cross_attn_layers):
                if cr_attn_type[pt_idx] == 'c':
                    pt_block = modules.PointTransformerBlock(
                        d_in=self.d_latent, d_hidden=self.d_latent, d_out=self.d_latent,
                        num_neighbors=cross_attn_neighbors, d_hidden_abstract=d_latent_local)
                elif cr_attn_type[pt_idx] == 's':
                    raise NotImplementedError()   # as the reference, line 252
                else:
                    raise ValueError()
                pt_blocks.append(pt_block)
                use_pt_inds.append(int((pt_idx + 1) * self.n_blocks / (cross_attn_layers + 1)))
            self.pt_blocks = nn.ModuleList(pt_blocks)
            self.use_pt_inds = {j: i for i, j in enumerate(use_pt_inds)}
        self._o4d_scene = None
        self._o4d_scene_key = None

    # ------------------------------------------------------------------ C-ABI plumbing
    def o4d_config(self):
        prec = ops.default_precision() if self.o4d_precision is None else int(self.o4d_precision)
        return ops.DecoderConfig(
            d_in=self.d_in, d_hidden=self.d_hidden, d_out=self.d_out, d_latent=self.d_latent,
            d_latent_local=self.d_latent_local, n_blocks=self.n_blocks,
            pos_encoding_freqs=self.pos_encoding_freqs, num_local_features=self.num_local_features,
            cross_attn_neighbors=self.cross_attn_neighbors, cross_attn_layers=self.cross_attn_layers,
            precision=prec)

    def o4d_params(self):
        """Parameter table in the order include/o4d.h documents for o4d_decoder_forward."""
        p = [self.lin_in.weight, self.lin_in.bias, self.lin_out.weight, self.lin_out.bias]
        for blk in self.blocks:
            p += [blk.fc_0.weight, blk.fc_0.bias, blk.fc_1.weight, blk.fc_1.bias]
        for lin in self.lin_z:
            p += [lin.weight, lin.bias]
        for blk in self.pt_blocks:
            p += blk.o4d_params()
        return p

    def o4d_scene(self, pcl_abstract, features_global):
        """Scene-constant state, rebuilt when the abstract cloud, the global embedding or any
        parameter changed (tensor identity + version counters).  When only the scene changed -- the next frame of a
        video, same weights, same abstract-cloud size -- the buffer is updated in place: everything that depends on the
        weights alone (composite Qa weights, packed tensor-core images, ...) is kept."""
        params = self.o4d_params()
        wkey = (tuple((p.data_ptr(), p._version) for p in params), self.o4d_precision, pcl_abstract.shape[0],
                pcl_abstract.device)
        key = (pcl_abstract.data_ptr(), pcl_abstract._version, tuple(pcl_abstract.shape),
               features_global.data_ptr(), features_global._version, wkey)
        if self._o4d_scene is None or key != self._o4d_scene_key:
            if self._o4d_scene is not None and self._o4d_scene_key[-1] == wkey:
                self._o4d_scene.update(params, pcl_abstract, features_global)
            else:
                self._o4d_scene = ops.DecoderScene(self.o4d_config(), params, pcl_abstract, features_global)
            self._o4d_scene_key = key
            # strong references: while cached, the allocator cannot hand the same addresses to a
            # different scene's tensors (which would make the identity key stale).
            self._o4d_scene_src = (pcl_abstract, features_global)
        return self._o4d_scene

    def forward(self, points_query, points_abstract, features_global, features_abstract, *extra):
        """points_query (B,N,4) | (N,4); points_abstract (B,M,3) or (B,M,3+E) with
        features_abstract None; features_global (B,D) -> (output (B,N,G), penult (B,N,H)).
        B must be 1 (reference line 317).  pipeline.py:193-194 passes one more positional flag and
        unpacks one more value: a trailing None is returned in that case."""
        if isinstance(self.activation, Swish):
            self.activation(points_query)
        if points_abstract is not None and features_abstract is None:
            pcl_abstract = points_abstract
        elif points_abstract is not None:
            pcl_abstract = torch.cat([points_abstract, features_abstract], dim=-1)
        else:
            pcl_abstract = None

        if len(points_query.shape) == 2:
            points_query = points_query.unsqueeze(0)
            pcl_abstract = pcl_abstract.unsqueeze(0) if pcl_abstract is not None else None
            features_global = features_global.unsqueeze(0)
            no_batch = True
        else:
            no_batch = False

        if self.num_local_features > 0:
            assert points_query.shape[0] == pcl_abstract.shape[0]
            assert points_query.shape[0] == features_global.shape[0]
            B = points_query.shape[0]
            assert B == 1, 'local attention mode takes one scene at a time (implicit.py:317), got shapes %s %s %s' % (
                tuple(points_query.shape), tuple(pcl_abstract.shape), tuple(features_global.shape))
            if self.local_mode == 'attention':
                assert points_query.shape[-1] == self.d_in
                assert features_global.shape[-1] + self.d_latent_local == self.d_latent
                assert pcl_abstract.shape[-1] == 3 + self.d_latent_local
                if _wants_grad(self, pcl_abstract, features_global):
                    from . import autograd
                    out, pen = autograd.decoder_train(self, ops._f32(points_query[0], 'points_query'),
                                                      ops._f32(pcl_abstract[0], 'points_abstract'),
                                                      ops._f32(features_global[0], 'features_global'))
                else:
                    scene = self.o4d_scene(pcl_abstract[0], features_global[0])
                    out, pen = ops.decoder_forward(self.o4d_config(), self.o4d_params(), scene, points_query[0])
                output, penult = out.unsqueeze(0), pen.unsqueeze(0)
            elif self.local_mode == 'feature':
                raise NotImplementedError("o4d: local_mode 'feature' is not on the released path")
            elif self.local_mode == 'function':
                raise NotImplementedError()   # as the reference, line 363
            else:
                raise ValueError()
        else:
            (output, penult) = super().do_forward(points_query, features_global)

        if no_batch:
            output = output.squeeze(0)
            penult = penult.squeeze(0)
        if extra:
            return (output, penult, None)
        return (output, penult)
