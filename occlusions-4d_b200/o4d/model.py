"""Host-side mirror of the reference's model/model.py: PointCompletionNetV3.

Same constructor kwargs (they are stored in checkpoints, train.py:345-347), same
parameter names / creation order / state_dict layout, same return tuple.  The whole
forward of one cloud is ONE call into libo4d.so (o4d_encoder_forward).
"""
import torch
from torch import nn

from . import modules
from . import ops
from .point_transformer_layer import _wants_grad


class PointCompletionNetV3(nn.Module):
    """Point-transformer encoder: decorated cloud -> abstract featurised cloud + global embedding."""

    def __init__(self, mixed_precision=False, n_input=4096, n_output=1024, d_in=6, d_out=6,
                 d_feat=32, down_blocks=3, up_blocks=2, transition_factor=4,
                 pt_num_neighbors=16, pt_norm_type='none', down_neighbors=8, abstract_levels=1,
                 skip_connections=False, enable_decoder=False, output_featurized=True,
                 output_global_emb=True, global_dim=512, fps_random_start=True):
        super().__init__()
        self.mixed_precision = mixed_precision
        self.n_input = n_input
        self.n_output = n_output
        self.d_in = d_in
        self.d_out = d_out
        self.d_feat = d_feat
        self.down_blocks = down_blocks
        self.up_blocks = up_blocks
        self.transition_factor = transition_factor
        self.pt_num_neighbors = pt_num_neighbors
        self.pt_norm_type = pt_norm_type
        self.down_neighbors = down_neighbors
        self.abstract_levels = abstract_levels
        self.skip_connections = skip_connections
        self.enable_decoder = enable_decoder
        self.output_featurized = output_featurized
        self.output_global_emb = output_global_emb
        self.global_dim = global_dim
        self.fps_random_start = fps_random_start

        if enable_decoder or skip_connections or not output_featurized or not output_global_emb:
            # reference lines 124-144 / 210-216: the point-cloud decoder (UpTransition) is dead in
            # both released configurations and references an undefined attribute (modules.py:288).
            raise NotImplementedError(
                'o4d supports the released encoder configuration only: enable_decoder=False, '
                'skip_connections=False, output_featurized=True, output_global_emb=True')
        if pt_norm_type not in ('none', 'layer'):
            raise NotImplementedError("o4d: pt_norm_type must be 'none' or 'layer'")

        # Creation order follows the reference (lines 75-146) so a seeded init reproduces it.
        dim = d_feat
        self.pre_mlp = nn.Sequential(nn.Linear(d_in, dim), nn.ReLU(), nn.Linear(dim, dim))
        blocks = []
        for _ in range(down_blocks):
            blocks.append(modules.PointTransformerBlock(
                d_in=dim, d_hidden=dim, d_out=dim, num_neighbors=pt_num_neighbors))
            blocks.append(modules.DownTransition(
                d_in=dim, d_out=dim * 2, factor=transition_factor, knn_k=down_neighbors,
                norm_type=pt_norm_type, fps_random_start=fps_random_start))
            dim *= 2
        blocks.append(modules.PointTransformerBlock(
            d_in=dim, d_hidden=dim, d_out=dim, num_neighbors=pt_num_neighbors))
        self.center_block_idx = len(blocks) - 1
        self.global_mlp = nn.Sequential(
            nn.Linear(dim, global_dim), nn.ReLU(), nn.Linear(global_dim, global_dim))
        if abstract_levels > 1:
            skip_mlps = []
            for level_idx in range(abstract_levels - 1):
                cur_dim = dim // int(2 ** (abstract_levels - 1 - level_idx))
                skip_mlps.append(nn.Linear(cur_dim, dim))
            self.abstract_skip_mlps = nn.ModuleList(skip_mlps)
        self.blocks = nn.ModuleList(blocks)
        self.o4d_precision = None

    # ------------------------------------------------------------------ C-ABI plumbing
    def o4d_config(self):
        prec = ops.default_precision() if self.o4d_precision is None else int(self.o4d_precision)
        return ops.EncoderConfig(
            d_in=self.d_in, d_feat=self.d_feat, down_blocks=self.down_blocks,
            transition_factor=self.transition_factor, pt_num_neighbors=self.pt_num_neighbors,
            down_neighbors=self.down_neighbors, norm=1 if self.pt_norm_type == 'layer' else 0,
            abstract_levels=self.abstract_levels, global_dim=self.global_dim, precision=prec)

    def o4d_params(self):
        """Parameter table in the order include/o4d.h documents for o4d_encoder_forward."""
        p = [self.pre_mlp[0].weight, self.pre_mlp[0].bias, self.pre_mlp[2].weight, self.pre_mlp[2].bias,
             self.global_mlp[0].weight, self.global_mlp[0].bias, self.global_mlp[2].weight,
             self.global_mlp[2].bias]
        if self.abstract_levels > 1:
            for lin in self.abstract_skip_mlps:
                p += [lin.weight, lin.bias]
        for blk in self.blocks:
            if isinstance(blk, modules.PointTransformerBlock):
                p += blk.o4d_params()
            else:
                q = blk.o4d_params()
                p += q if self.pt_norm_type == 'layer' else q[:2]
        return p

    def forward(self, pcl, return_intermediate, *extra):
        """pcl (B,N,d_in) -> (pcl_out (B,M,3+E), x_global (B,F), layer_coords | None).

        pipeline.py:93-94 passes one more positional flag and unpacks one more value than the
        published model.py:148 defines; when that flag is present a trailing None is returned
        so both callers run unmodified."""
        assert pcl.dim() == 3 and pcl.shape[-1] == self.d_in
        cfg = self.o4d_config()
        params = self.o4d_params()
        B, N, _ = pcl.shape
        outs, globs, coords = [], [], []
        train_path = _wants_grad(self, pcl)
        for b in range(B):
            starts = None
            if self.fps_random_start:
                starts, nl = [], N
                for _ in range(self.down_blocks):
                    starts.append(int(torch.randint(0, nl, (1,))))
                    nl = -(-nl // self.transition_factor)
            if train_path:
                from . import autograd
                res = autograd.encoder_train(self, ops._f32(pcl[b], 'pcl').contiguous(), starts)
            else:
                res = ops.encoder_forward(cfg, params, pcl[b], starts, return_levels=bool(return_intermediate))
            outs.append(res[0])
            globs.append(res[1])
            if return_intermediate:
                coords.append(res[2])
        pcl_out = torch.stack(outs)
        x_global = torch.stack(globs)
        layer_coords = None
        if return_intermediate:
            # reference lines 161-199: input coords, pos0, then the coords after every block.
            per_level = [torch.stack([c[l] for c in coords]) for l in range(self.down_blocks + 1)]
            layer_coords = [per_level[0], per_level[0]]
            for l in range(self.down_blocks):
                layer_coords.append(per_level[l])       # after the PT block of level l
                layer_coords.append(per_level[l + 1])   # after the down transition
            layer_coords.append(per_level[self.down_blocks])  # after the centre block
        if extra:
            return (pcl_out, x_global, layer_coords, None)
        return (pcl_out, x_global, layer_coords)
