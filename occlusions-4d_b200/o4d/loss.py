"""Objective functions of the implicit heads (host-side mirror of the reference's loss.py).

`MyLosses` keeps the reference's constructor, `per_example` and `entire_batch` (loss.py:15-294); the four
per-frame heads are ONE fused device call (`implicit_loss_heads`: o4d_implicit_loss_forward_f32, gradients by
o4d_implicit_loss_backward_f32) instead of ~80 small kernels and a host sync per boolean mask.  CUDA only.
"""

import torch

from . import _lib, ops

COLOR_MODES = {'rgb': 0, 'rgb_nosigmoid': 0, 'hsv': 1, 'bins': 2}


def get_track_idx(color_mode):
    """Output column of the tracking logit (utils/utils.py:204-224)."""
    return {'rgb': 4, 'rgb_nosigmoid': 4, 'hsv': 15, 'bins': 10}[color_mode]


class _ImplicitLossHeads(torch.autograd.Function):
    """(output (n, g), target (n, 6)) -> losses (4,) = (rgb, dens, segm, track); loss.py:50-194."""

    @staticmethod
    def forward(ctx, output, target, color_mode, semantic_classes, track_idx):
        out2, ldo = ops._rows(output, 'implicit_output')
        tgt2, ldt = ops._rows(target, 'implicit_target')
        assert tgt2.shape[0] == out2.shape[0] and tgt2.shape[1] == 6, 'implicit_target must be (n, 6)'
        n, g = out2.shape
        L = _lib.lib()
        with torch.cuda.device(out2.device):
            losses = torch.empty((4,), dtype=torch.float32, device=out2.device)
            stats = torch.empty((12,), dtype=torch.float64, device=out2.device)
            ws = ops.workspace(out2.device, L.o4d_implicit_loss_workspace_bytes(n), slot=5)
            rc = L.o4d_implicit_loss_forward_f32(
                ops._ptr(out2), n, g, ldo, ops._ptr(tgt2), ldt, int(color_mode), int(semantic_classes),
                int(track_idx), ops._ptr(losses), ops._ptr(stats), ops._ptr(ws), ws.numel(), ops._stream(out2))
        _lib.check(rc, 'o4d_implicit_loss_forward_f32')
        ctx.save_for_backward(out2, tgt2, stats)
        ctx.meta = (ldo, ldt, int(color_mode), int(semantic_classes), int(track_idx), output.shape, output.dtype)
        return losses

    @staticmethod
    def backward(ctx, dlosses):
        out2, tgt2, stats = ctx.saved_tensors
        ldo, ldt, color_mode, semantic_classes, track_idx, shape, dtype = ctx.meta
        n, g = out2.shape
        w = dlosses.detach().float().contiguous()
        with torch.cuda.device(out2.device):
            dout = torch.empty((n, g), dtype=torch.float32, device=out2.device)
            rc = _lib.lib().o4d_implicit_loss_backward_f32(
                ops._ptr(out2), n, g, ldo, ops._ptr(tgt2), ldt, color_mode, semantic_classes, track_idx,
                ops._ptr(stats), ops._ptr(w), ops._ptr(dout), g, ops._stream(out2))
        _lib.check(rc, 'o4d_implicit_loss_backward_f32')
        return dout.reshape(shape).to(dtype), None, None, None, None


def implicit_loss_heads(implicit_output, implicit_target, color_mode='rgb', semantic_classes=0, track=True):
    """All four heads of one frame in one pass -> (4,) device tensor (loss_rgb, loss_dens, loss_segm,
    loss_track); disabled heads read 0.  Inputs (..., n, g) / (..., n, 6) are flattened over leading dims."""
    g = implicit_output.shape[-1]
    out2 = implicit_output.reshape(-1, g)
    tgt2 = implicit_target.reshape(-1, implicit_target.shape[-1])
    ti = get_track_idx(color_mode) if track else -1
    if ti >= g:
        ti = -1
    return _ImplicitLossHeads.apply(out2, tgt2, COLOR_MODES[color_mode], int(semantic_classes), ti)


class MyLosses():
    """Same constructor and methods as the reference's MyLosses (loss.py:15-294)."""

    def __init__(self, stage, logger, mixed_precision, color_lw, density_lw, segmentation_lw,
                 tracking_lw, color_mode, semantic_classes, past_frames, future_frames):
        self.stage = stage
        self.logger = logger
        self.mixed_precision = mixed_precision
        self.color_lw = color_lw
        self.density_lw = density_lw
        self.segmentation_lw = segmentation_lw
        self.tracking_lw = tracking_lw
        self.color_mode = color_mode
        self.semantic_classes = semantic_classes
        self.past_frames = past_frames
        self.future_frames = future_frames

    def _heads(self, implicit_output, implicit_target, segm=True, track=True):
        return implicit_loss_heads(implicit_output, implicit_target, self.color_mode,
                                   self.semantic_classes if segm else 0, track)

    # the single-head entry points of the reference (each runs the fused pass and picks its scalar)
    def implicit_density_loss(self, implicit_output, implicit_target):
        return self._heads(implicit_output, implicit_target, segm=False, track=False)[1]

    def implicit_color_loss(self, implicit_output, implicit_target):
        return self._heads(implicit_output, implicit_target, segm=False, track=False)[0]

    def implicit_segm_loss(self, implicit_output, implicit_target):
        return self._heads(implicit_output, implicit_target, segm=True, track=False)[2]

    def implicit_track_loss(self, implicit_output, implicit_target):
        return self._heads(implicit_output, implicit_target, segm=False, track=True)[3]

    def per_example(self, pcl_target, pcl_target_size, implicit_output, implicit_target):
        """loss.py:196-253: per (example, frame) heads averaged within this GPU -> (loss_rgb, loss_dens,
        loss_segm, loss_track), None for heads whose weight is zero.  One fused call per (example, frame)."""
        (B, M, E) = pcl_target[0].shape
        assert torch.all(torch.as_tensor(pcl_target_size) <= M)
        assert implicit_output is not None
        use_segm = self.segmentation_lw > 0.0
        use_track = self.tracking_lw > 0.0
        rows = []
        for i in range(B):
            for time_idx in range(self.past_frames + self.future_frames):
                rows.append(self._heads(implicit_output[time_idx][i:i + 1], implicit_target[time_idx][i:i + 1],
                                        segm=use_segm, track=use_track))
        mean = torch.stack(rows).mean(dim=0)
        return (mean[0] if self.color_lw > 0.0 else None, mean[1] if self.density_lw > 0.0 else None,
                mean[2] if use_segm else None, mean[3] if use_track else None)

    def entire_batch(self, total_step, loss_rgb, loss_dens, loss_segm, loss_track, points_query,
                     implicit_output, features_global):
        """loss.py:255-294: average over GPUs, weight, report."""
        loss_rgb = loss_rgb.mean() if torch.is_tensor(loss_rgb) else 0.0
        loss_dens = loss_dens.mean() if torch.is_tensor(loss_dens) else 0.0
        loss_segm = loss_segm.mean() if torch.is_tensor(loss_segm) else 0.0
        loss_track = loss_track.mean() if torch.is_tensor(loss_track) else 0.0
        total_loss = loss_rgb * self.color_lw + loss_dens * self.density_lw + \
            loss_segm * self.segmentation_lw + loss_track * self.tracking_lw
        report = (lambda k, v: self.logger.report_scalar(self.stage + k, v, remember=True)) \
            if self.logger is not None else (lambda k, v: None)
        report('/total_loss', total_loss.item())
        scalars = []
        for key, val in (('/loss_rgb', loss_rgb), ('/loss_dens', loss_dens), ('/loss_segm', loss_segm),
                         ('/loss_track', loss_track)):
            if torch.is_tensor(val) or val != 0.0:
                val = val.item() if torch.is_tensor(val) else val
                report(key, val)
            scalars.append(val)
        return (total_loss,) + tuple(scalars)
