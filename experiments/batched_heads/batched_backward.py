"""Backward of the appearance heads for H heads at once (experiments; round-2 candidate for tensoir_b200/heads.py).

Today each head runs its own chain in heads._head_backward: ~45 small launches per head, 4 heads per step = the largest
block of "glue" in profiles/r1_step_phase_attribution.txt.  With equal sample counts per head (the three heads of the
primary group and the BRDF head at the jittered points all see the same appearance list) the chains can be stacked:
batched GEMMs for dgrad / wgrad, stacked elementwise ops, one shared basis wgrad — H x fewer launches and 4 x larger
GEMM batches.  To avoid stacking copies the heads' forward launches would dump straight into [H, n, *] buffers.

Pure torch; pinned against the per-head chain on the CPU by test_batched_backward.py.
"""
import torch

from tensoir_b200.heads import _wgrad


def _wgrad_batched(g, a, splits=8):
    """Per-head g_h^T @ a_h for stacked operands g [H,n,M], a [H,n,K] -> [H,M,K]; reduction over n split like _wgrad."""
    H, n, M = g.shape
    per = n // splits
    if per < 64:
        return torch.bmm(g.transpose(1, 2), a)
    main = per * splits
    out = torch.bmm(g[:, :main].reshape(H * splits, per, M).transpose(1, 2),
                    a[:, :main].reshape(H * splits, per, -1)).view(H, splits, M, -1).sum(1)
    if main < n:
        out = out + torch.bmm(g[:, main:].transpose(1, 2), a[:, main:])
    return out


def heads_backward_batched(is_tanh, g_out, out, xl, inp, h1, h2, w0, w1, w2, basis):
    """is_tanh [H] bool; g_out / out [H,n,4] (heads with 3 outputs zero-padded); xl [H,n,3*aC]; inp [H,n,in_dim];
    h1, h2 [H,n,hid]; w0 [H,hid,in_dim]; w1 [H,hid,hid]; w2 [H,4,hid] (zero-padded rows); basis [F,3*aC].
    -> gw0 [H,hid,in_dim], gb0 [H,hid], gw1, gb1, gw2 [H,4,hid], gb2 [H,4], gbasis [F,3*aC] (summed over heads),
       gxl [H,n,3*aC] (gradient w.r.t. the light-scaled products; the per-head light factor is applied by the caller)."""
    H, n, _ = out.shape
    dact = torch.where(is_tanh.view(H, 1, 1), 1 - out * out, out * (1 - out))
    gz3 = g_out * dact
    gw2, gb2 = torch.bmm(gz3.transpose(1, 2), h2), gz3.sum(1)
    gz2 = torch.bmm(gz3, w2) * (h2 > 0)
    gw1, gb1 = _wgrad_batched(gz2, h1), gz2.sum(1)
    gz1 = torch.bmm(gz2, w1) * (h1 > 0)
    gw0, gb0 = _wgrad_batched(gz1, inp), gz1.sum(1)
    gin = torch.bmm(gz1, w0)                                   # [H, n, in_dim]
    F, pe = basis.shape[0], 2
    s0, c0 = F + 3, F + 3 + F * pe
    freqs = 2.0 ** torch.arange(pe, device=out.device, dtype=inp.dtype)
    gs, gc = gin[..., s0:s0 + F * pe].reshape(H, n, F, pe), gin[..., c0:c0 + F * pe].reshape(H, n, F, pe)
    sn, cs = inp[..., s0:s0 + F * pe].reshape(H, n, F, pe), inp[..., c0:c0 + F * pe].reshape(H, n, F, pe)
    gfeat = gin[..., :F] + ((gs * cs - gc * sn) * freqs).sum(-1)           # [H, n, F]
    gbasis = _wgrad(gfeat.reshape(H * n, F), xl.reshape(H * n, -1))
    gxl = (gfeat.reshape(H * n, F) @ basis).view(H, n, -1)
    return gw0, gb0, gw1, gb1, gw2, gb2, gbasis, gxl
