"""Total-variation regulariser of the VM planes as two kernel launches.

Reference: ``TVLoss`` (utils.py:143-162) summed by ``TensorVMSplit.TV_loss_density / TV_loss_app``
(tensoRF_rotated_lights.py:80-92), used by the radiance-only phase of training (train_tensoIR.py:276-285, weights decayed
every iteration, switched off when relighting starts :398-399).  The reference's slice / pow / sum graph is ~14 launches per
plane and direction of autograd; here the three planes of one call are one forward launch (the value) and one backward
launch that adds the 5-point stencil straight into the gradient buffers (`.grad` itself when it is a dense tensor of the
parameter's layout, like primary_fused does).
"""
from __future__ import annotations

import torch

from . import _lib


def _layout(p):
    """-> channel_last flag of a dense [1,C,H,W] parameter (None: neither layout)."""
    if p.dim() != 4 or p.shape[0] != 1 or p.dtype != torch.float32:
        return None
    if p.is_contiguous():                     # also true for C == 1 or H*W == 1 channel-last tensors: same bytes
        return 0
    if p.is_contiguous(memory_format=torch.channels_last):
        return 1
    return None


def _table(planes, weight, grads):
    tab = (_lib.TirTvPlane * len(planes))()
    for k, p in enumerate(planes):
        _, Cc, H, W = p.shape
        count_h, count_w = Cc * (H - 1) * W, Cc * H * (W - 1)
        t = tab[k]
        t.x = p.data_ptr()
        t.grad = None if grads is None else grads[k].data_ptr()
        t.H, t.W, t.C, t.channel_last = H, W, Cc, _layout(p)
        # 2 * TVLoss_weight * (. / count) / batch_size, times the 1e-2 of TV_loss_*; a zero count is the reference's 0/0
        t.scale_h = 2.0 * weight / count_h if count_h else float("nan")
        t.scale_w = 2.0 * weight / count_w if count_w else float("nan")
    return tab


class _TV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, inplace, *planes):
        lib = _lib.load()
        ctx.weight, ctx.inplace = weight, inplace
        ctx.save_for_backward(*planes)
        out = torch.empty(1, device=planes[0].device, dtype=torch.float32)
        tab = _table([p.detach() for p in planes], weight, None)
        _lib.check(lib.tir_tv_loss(tab, len(planes), _lib.dptr(out), _lib.stream_ptr()), "tir_tv_loss")
        return out[0]

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        planes = ctx.saved_tensors
        bufs, rets = [], []
        for p, need in zip(planes, ctx.needs_input_grad[2:]):
            g = getattr(p, "grad", None)
            if (ctx.inplace and need and g is not None and g.dtype == torch.float32 and g.shape == p.shape
                    and all(a == b for a, b, n in zip(g.stride(), p.stride(), p.shape) if n > 1)):
                bufs.append(g)
                rets.append(None)             # accumulated in place: autograd gets nothing to add
            else:
                z = torch.zeros_like(p, memory_format=torch.preserve_format)
                bufs.append(z)
                rets.append(z if need else None)
        go = gout.detach().reshape(1).float().contiguous()
        tab = _table([p.detach() for p in planes], ctx.weight, bufs)
        _lib.check(lib.tir_tv_loss_bwd(tab, len(planes), _lib.dptr(go), _lib.stream_ptr()), "tir_tv_loss_bwd")
        return (None, None, *rets)


def tv_planes(planes, reg, factor=1e-2, inplace=True):
    """``sum(reg(p) * factor for p in planes)`` — on the device in one launch when ``reg`` is the reference's TVLoss
    (recognised by its ``TVLoss_weight`` attribute) and the planes are CUDA parameters; any other callable is simply
    called (that is the caller's own regulariser, not a fallback of this kernel)."""
    planes = list(planes)
    w = getattr(reg, "TVLoss_weight", None)
    if (w is None or not planes or not all(p.is_cuda and _layout(p) is not None for p in planes)
            or len(planes) > _lib.TV_MAX_PLANES):
        return sum(reg(p) * factor for p in planes)
    return _TV.apply(float(w) * factor, bool(inplace), *planes)
