"""Differentiable wrapper of the fused shading kernels (csrc/tir_shade.cu): the quadrature of the rendering
equation at the surface points, models/relight_utils.py:452-475, forward and analytic backward."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class _Shade(torch.autograd.Function):
    @staticmethod
    def forward(ctx, normal, albedo, rough3, fresnel, direct, view, light_idx, dirs, weight, vis, indirect):
        lib = _lib.load()
        ts = [t.detach().float().contiguous() for t in (normal, albedo, rough3, fresnel, direct, view, dirs, weight,
                                                         vis, indirect)]
        normal, albedo, rough3, fresnel, direct, view, dirs, weight, vis, indirect = ts
        li = light_idx.detach().reshape(-1).to(torch.int32).contiguous()
        bs, nl = normal.shape[0], dirs.shape[0]
        rgb = torch.zeros(bs, 3, device=normal.device)
        _lib.check(lib.tir_shade_fwd(_lib.dptr(normal), _lib.dptr(albedo), _lib.dptr(rough3), _lib.dptr(fresnel),
                                     _lib.dptr(view), _lib.dptr(li, torch.int32), bs, _lib.dptr(dirs),
                                     _lib.dptr(weight), nl, _lib.dptr(direct), direct.shape[0], _lib.dptr(vis),
                                     _lib.dptr(indirect), _lib.dptr(rgb), _lib.stream_ptr()), "tir_shade_fwd")
        ctx.save_for_backward(normal, albedo, rough3, fresnel, direct, view, li, dirs, weight, vis, indirect)
        return rgb

    @staticmethod
    def backward(ctx, g_rgb):
        normal, albedo, rough3, fresnel, direct, view, li, dirs, weight, vis, indirect = ctx.saved_tensors
        lib = _lib.load()
        g_rgb = g_rgb.contiguous().float()
        bs, nl = normal.shape[0], dirs.shape[0]
        gn, ga, gr, gf = (torch.zeros_like(normal), torch.zeros_like(albedo), torch.zeros_like(rough3),
                          torch.zeros_like(fresnel))
        gd = torch.zeros_like(direct)
        _lib.check(lib.tir_shade_bwd(_lib.dptr(normal), _lib.dptr(albedo), _lib.dptr(rough3), _lib.dptr(fresnel),
                                     _lib.dptr(view), _lib.dptr(li, torch.int32), bs, _lib.dptr(dirs),
                                     _lib.dptr(weight), nl, _lib.dptr(direct), direct.shape[0], _lib.dptr(vis),
                                     _lib.dptr(indirect), _lib.dptr(g_rgb), _lib.dptr(gn), _lib.dptr(ga),
                                     _lib.dptr(gr), _lib.dptr(gf), _lib.dptr(gd), _lib.stream_ptr()), "tir_shade_bwd")
        return gn, ga, gr, gf, gd, None, None, None, None, None, None


def shade(normal, albedo, rough3, fresnel, direct, view, light_idx, dirs, weight, vis, indirect):
    """sum_l (albedo/pi + GGX) * (vis * direct[light] + indirect) * clamp(l.n, 0) * weight_l  -> [bs, 3] (linear).
    Differentiable w.r.t. normal, albedo, rough3, fresnel and direct ([n_lights, n_dirs, 3])."""
    nl = dirs.shape[0]
    return _Shade.apply(normal, albedo, rough3, fresnel, direct, view, light_idx, dirs, weight,
                        vis.reshape(-1, nl), indirect.reshape(-1, nl, 3))
