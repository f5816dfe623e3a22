"""Differentiable wrapper of the fused shading kernels (csrc/tir_shade.cu): the quadrature of the rendering
equation at the surface points, models/relight_utils.py:452-475, forward and analytic backward."""
from __future__ import annotations

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


class _ShadeHits(torch.autograd.Function):
    """Whole ray batch masked by acc_mask: secondary rays + quadrature + clip + sRGB, no list of hits."""

    @staticmethod
    def forward(ctx, normal, albedo, rough1, fresnel, direct, rays, mask, light_idx, dirs, weight, vis, indirect, srgb):
        lib = _lib.load()
        normal, albedo, rough1, fresnel, direct, rays, dirs, weight, vis, indirect = (
            t.detach().float().contiguous() for t in (normal, albedo, rough1, fresnel, direct, rays, dirs, weight, vis,
                                                      indirect))
        li = light_idx.detach().reshape(-1).to(torch.int32).contiguous()
        mask = mask.detach().contiguous()
        n, nl = normal.shape[0], dirs.shape[0]
        rgb = torch.empty(n, 3, device=normal.device)
        lin = torch.empty(n, 3, device=normal.device)
        _lib.check(lib.tir_shade_hits_fwd(_lib.dptr(rays), _lib.dptr(mask, torch.bool), _lib.dptr(normal),
                                          _lib.dptr(albedo), _lib.dptr(rough1), _lib.dptr(fresnel),
                                          _lib.dptr(li, torch.int32), n, _lib.dptr(dirs), _lib.dptr(weight), nl,
                                          _lib.dptr(direct), direct.shape[0], _lib.dptr(vis), _lib.dptr(indirect),
                                          int(bool(srgb)), _lib.dptr(rgb), _lib.dptr(lin), _lib.stream_ptr()),
                   "tir_shade_hits_fwd")
        ctx.save_for_backward(normal, albedo, rough1, fresnel, direct, rays, mask, li, dirs, weight, vis, indirect, lin)
        ctx.srgb = int(bool(srgb))
        return rgb

    @staticmethod
    def backward(ctx, g_rgb):
        normal, albedo, rough1, fresnel, direct, rays, mask, li, dirs, weight, vis, indirect, lin = ctx.saved_tensors
        lib = _lib.load()
        g_rgb = g_rgb.contiguous().float()
        n, nl = normal.shape[0], dirs.shape[0]
        gn, ga, gr, gf = (torch.empty_like(normal), torch.empty_like(albedo), torch.empty_like(rough1),
                          torch.empty_like(fresnel))
        gd = torch.zeros_like(direct)
        _lib.check(lib.tir_shade_hits_bwd(_lib.dptr(rays), _lib.dptr(mask, torch.bool), _lib.dptr(normal),
                                          _lib.dptr(albedo), _lib.dptr(rough1), _lib.dptr(fresnel),
                                          _lib.dptr(li, torch.int32), n, _lib.dptr(dirs), _lib.dptr(weight), nl,
                                          _lib.dptr(direct), direct.shape[0], _lib.dptr(vis), _lib.dptr(indirect),
                                          ctx.srgb, _lib.dptr(lin), _lib.dptr(g_rgb), _lib.dptr(gn), _lib.dptr(ga),
                                          _lib.dptr(gr), _lib.dptr(gf), _lib.dptr(gd), _lib.stream_ptr()),
                   "tir_shade_hits_bwd")
        return gn, ga, gr, gf, gd, None, None, None, None, None, None, None, None


def shade_hits(normal, albedo, rough1, fresnel, direct, rays, mask, light_idx, dirs, weight, vis, indirect, srgb=True):
    """rgb_with_brdf for a whole ray batch: rows with ``mask`` are shaded (quadrature, clip to [0,1], sRGB), the others
    are 1.  Differentiable w.r.t. normal, albedo, rough1 ([n,1]), fresnel and direct."""
    nl = dirs.shape[0]
    return _ShadeHits.apply(normal, albedo, rough1, fresnel, direct, rays, mask, light_idx, dirs, weight,
                            vis.reshape(-1, nl), indirect.reshape(-1, nl, 3), srgb)
