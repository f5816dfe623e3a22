"""Differentiable wrappers of the fused end of the primary march (csrc/tir_tail.cu): the per-sample tail between the
appearance heads and the per-ray sums, and the per-ray epilogue of TensorBase.forward
(tensorBase_rotated_lights.py:930-1036).  Two launches forward, two backward, instead of ~200 small torch kernels."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

TAIL_CHANNELS = 14


def _c(t):
    return None if t is None else t.detach().float().contiguous()


class _Tail(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, ray, rgb, brdf, brdfj, vn, dn, viewdirs, n_rays):
        lib = _lib.load()
        w, rgb, brdf, brdfj, vn, dn, viewdirs = (_c(t) for t in (w, rgb, brdf, brdfj, vn, dn, viewdirs))
        ray = ray.detach().to(torch.int64).contiguous()
        packed = torch.zeros(n_rays, TAIL_CHANNELS, device=w.device)
        _lib.check(lib.tir_tail_fwd(w.shape[0], _lib.dptr(w), _lib.dptr(ray, torch.int64), _lib.dptr(rgb),
                                    _lib.dptr(brdf), _lib.dptr(brdfj), _lib.dptr(vn), _lib.dptr(dn, allow_none=True),
                                    _lib.dptr(viewdirs), _lib.dptr(packed), _lib.stream_ptr()), "tir_tail_fwd")
        ctx.has_dn = dn is not None
        ctx.save_for_backward(w, ray, rgb, brdf, brdfj, vn, viewdirs, *([dn] if dn is not None else []))
        return packed

    @staticmethod
    def backward(ctx, g_packed):
        w, ray, rgb, brdf, brdfj, vn, viewdirs, *rest = ctx.saved_tensors
        dn = rest[0] if ctx.has_dn else None
        lib = _lib.load()
        g_w, g_rgb, g_brdf, g_brdfj, g_vn = (torch.empty_like(t) for t in (w, rgb, brdf, brdfj, vn))
        g_dn = torch.empty_like(dn) if dn is not None else None
        _lib.check(lib.tir_tail_bwd(w.shape[0], _lib.dptr(w), _lib.dptr(ray, torch.int64), _lib.dptr(rgb),
                                    _lib.dptr(brdf), _lib.dptr(brdfj), _lib.dptr(vn), _lib.dptr(dn, allow_none=True),
                                    _lib.dptr(viewdirs), _lib.dptr(g_packed.contiguous().float()), _lib.dptr(g_w),
                                    _lib.dptr(g_rgb), _lib.dptr(g_brdf), _lib.dptr(g_brdfj), _lib.dptr(g_vn),
                                    _lib.dptr(g_dn, allow_none=True), _lib.stream_ptr()), "tir_tail_bwd")
        return g_w, None, g_rgb, g_brdf, g_brdfj, g_vn, g_dn, None, None


def fused_tail(w, ray, rgb, brdf, brdfj, vn, dn, viewdirs, n_rays):
    """-> [n_rays, 14] = per-ray sums of w * [rgb, vn, albedo, roughness, albedo cost, roughness cost, |vn - dn|^2,
    clamp(viewdir . vn, 0)] (the last two channels are 0 when ``dn`` is None)."""
    if w.shape[0] == 0:
        return torch.zeros(n_rays, TAIL_CHANNELS, device=w.device)
    return _Tail.apply(w, ray, rgb, brdf, brdfj, vn, dn, viewdirs, int(n_rays))


def _maps(ts):
    m = _lib.TirRayMaps()
    for k, t in zip(("rgb", "depth", "normal", "albedo", "rough", "fresnel", "nd", "no"), ts):
        m_ptr = None if t is None else t.data_ptr()
        setattr(m, k, m_ptr)
    return m


class _Epilogue(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed, acc, depth, rays, fresnel0, bg):
        lib = _lib.load()
        packed, acc, depth, rays = (_c(t) for t in (packed, acc, depth, rays))
        n, dev = packed.shape[0], packed.device
        outs = [torch.empty(n, 3, device=dev), torch.empty(n, device=dev), torch.empty(n, 3, device=dev),
                torch.empty(n, 3, device=dev), torch.empty(n, 1, device=dev), torch.empty(n, 3, device=dev),
                torch.empty(n, 1, device=dev), torch.empty(n, 1, device=dev)]
        mask = torch.empty(n, dtype=torch.bool, device=dev)          # one byte per ray, written as 0 / 1
        losses = torch.zeros(2, device=dev)
        maps = _maps(outs)
        _lib.check(lib.tir_epilogue_fwd(n, _lib.dptr(packed), _lib.dptr(acc), _lib.dptr(depth), _lib.dptr(rays),
                                        float(fresnel0), int(bool(bg)), C.byref(maps), _lib.dptr(mask, torch.bool),
                                        _lib.dptr(losses), _lib.stream_ptr()), "tir_epilogue_fwd")
        ctx.save_for_backward(packed, acc, depth, rays)
        ctx.consts = (float(fresnel0), int(bool(bg)))
        ctx.mark_non_differentiable(mask)
        loss_albedo, loss_rough = losses[0], losses[1]
        return (*outs, mask, loss_albedo, loss_rough)

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_normal, g_albedo, g_rough, g_fresnel, g_nd, g_no, _g_mask, g_la, g_lr):
        packed, acc, depth, rays = ctx.saved_tensors
        lib = _lib.load()
        n = packed.shape[0]
        gs = [None if g is None else g.contiguous().float()
              for g in (g_rgb, g_depth, g_normal, g_albedo, g_rough, g_fresnel, g_nd, g_no)]
        g_la = None if g_la is None else g_la.reshape(1).contiguous().float()
        g_lr = None if g_lr is None else g_lr.reshape(1).contiguous().float()
        maps = _maps(gs)
        g_packed, g_acc, g_dep = torch.empty_like(packed), torch.empty_like(acc), torch.empty_like(depth)
        fresnel0, bg = ctx.consts
        _lib.check(lib.tir_epilogue_bwd(n, _lib.dptr(packed), _lib.dptr(acc), _lib.dptr(depth), _lib.dptr(rays),
                                        fresnel0, bg, C.byref(maps), _lib.dptr(g_la, allow_none=True),
                                        _lib.dptr(g_lr, allow_none=True), _lib.dptr(g_packed), _lib.dptr(g_acc),
                                        _lib.dptr(g_dep), _lib.stream_ptr()), "tir_epilogue_bwd")
        return g_packed, g_acc, g_dep, None, None, None


def epilogue(packed, acc, depth, rays, fresnel0, bg):
    """Per-ray end of TensorBase.forward -> (rgb_map [n,3] sRGB, depth_map [n], normal_map [n,3], albedo_map [n,3],
    roughness_map [n,1], fresnel_map [n,3], normals_diff_map [n,1], normals_orientation_loss_map [n,1],
    acc_mask [n] bool, albedo_smoothness_loss, roughness_smoothness_loss)."""
    return _Epilogue.apply(packed, acc, depth, rays, fresnel0, bool(bg))
