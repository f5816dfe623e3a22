"""torch.autograd.Function wrappers around the point-wise VM kernels (csrc/tir_vm.cu).

PyTorch owns autograd; the forward gathers and the backward scatters (the reference's
grid_sampler_2d / grid_sampler_2d_backward launches) run in the CUDA
kernels over channel-last shadows, and gradients are handed back in the parameters' NCHW layout.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, ops


def _grad_shadows(df, kind):
    planes = df.dplane if kind == "density" else df.aplane
    lines = df.dline if kind == "density" else df.aline
    gp = [torch.zeros_like(p) for p in planes]
    gl = [torch.zeros_like(l) for l in lines]
    return gp, gl


def _ptr_array(ts):
    arr = (C.c_void_p * 3)()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr()
    return arr


def _to_param_layout(gp, gl):
    """channel-last [H,W,C] / [D,1,C] grads -> NCHW views [1,C,H,W] / [1,C,D,1]."""
    out = [g.permute(2, 0, 1).unsqueeze(0) for g in gp]
    out += [g.permute(2, 0, 1).unsqueeze(0) for g in gl]
    return out


def _needs_grad(ts):
    return torch.is_grad_enabled() and any(t.requires_grad for t in ts)


class _DensityFeature(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, xn, *params):
        feat, _ = ops.density_points(model, xn)
        ctx.model = model
        ctx.save_for_backward(xn)
        return feat

    @staticmethod
    def backward(ctx, g):
        model = ctx.model
        (xn,) = ctx.saved_tensors
        lib = _lib.load()
        df = ops.device_field(model)
        f = df.refresh(model)
        gp, gl = _grad_shadows(df, "density")
        g = g.contiguous().float()
        _lib.check(lib.tir_vm_density_bwd(C.byref(f), _lib.dptr(xn), xn.shape[0], _lib.dptr(g), _ptr_array(gp),
                                          _ptr_array(gl), _lib.stream_ptr()), "tir_vm_density_bwd")
        return (None, None, *_to_param_layout(gp, gl))


def density_feature(model, xn):
    """compute_densityfeature (tensoRF_rotated_lights.py:95-110), differentiable w.r.t. the density factors."""
    xn = xn.detach().reshape(-1, 3).float().contiguous()
    params = list(model.density_plane) + list(model.density_line)
    if _needs_grad(params):
        return _DensityFeature.apply(model, xn, *params)
    return ops.density_points(model, xn)[0]


class _DensityFeatureGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, xn, *params):
        lib = _lib.load()
        f = ops.device_field(model).refresh(model)
        n = xn.shape[0]
        feat = torch.empty(n, device=xn.device)
        dfdx = torch.empty(n, 3, device=xn.device)
        _lib.check(lib.tir_vm_density_grad(C.byref(f), _lib.dptr(xn), n, _lib.dptr(feat), _lib.dptr(dfdx),
                                           _lib.stream_ptr()), "tir_vm_density_grad")
        ctx.model = model
        ctx.save_for_backward(xn)
        return feat, dfdx

    @staticmethod
    def backward(ctx, g_feat, g_dfdx):
        model = ctx.model
        (xn,) = ctx.saved_tensors
        lib = _lib.load()
        df = ops.device_field(model)
        f = df.refresh(model)
        gp, gl = _grad_shadows(df, "density")
        gf = None if g_feat is None else g_feat.contiguous().float()
        gd = None if g_dfdx is None else g_dfdx.contiguous().float()
        _lib.check(lib.tir_vm_density_grad_bwd(C.byref(f), _lib.dptr(xn), xn.shape[0],
                                               None if gf is None else _lib.dptr(gf),
                                               None if gd is None else _lib.dptr(gd), _ptr_array(gp), _ptr_array(gl),
                                               _lib.stream_ptr()), "tir_vm_density_grad_bwd")
        return (None, None, *_to_param_layout(gp, gl))


def density_feature_and_grad(model, xn):
    """(feature [n], d feature / d x_hat [n,3]) with the clamped-index sampler of relight_utils.py:57-107
    (compute_densityfeature_with_xyz_grad, tensoRF_rotated_lights.py:113-129); differentiable w.r.t. the
    density factors in BOTH outputs (the reference's create_graph=True double backward, tensorBase:846-853)."""
    xn = xn.detach().reshape(-1, 3).float().contiguous()
    params = list(model.density_plane) + list(model.density_line)
    if _needs_grad(params):
        return _DensityFeatureGrad.apply(model, xn, *params)
    with torch.no_grad():
        return _DensityFeatureGrad.apply(model, xn, *params)


class _AppProducts(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, xn, *params):
        lib = _lib.load()
        f = ops.device_field(model).refresh(model)
        n = xn.shape[0]
        out = torch.empty(n, 3 * f.aC, device=xn.device)
        _lib.check(lib.tir_vm_app_products(C.byref(f), _lib.dptr(xn), n, _lib.dptr(out), _lib.stream_ptr()),
                   "tir_vm_app_products")
        ctx.model = model
        ctx.save_for_backward(xn)
        return out

    @staticmethod
    def backward(ctx, g):
        model = ctx.model
        (xn,) = ctx.saved_tensors
        lib = _lib.load()
        df = ops.device_field(model)
        f = df.refresh(model)
        gp, gl = _grad_shadows(df, "app")
        g = g.contiguous().float()
        _lib.check(lib.tir_vm_app_products_bwd(C.byref(f), _lib.dptr(xn), xn.shape[0], _lib.dptr(g), _ptr_array(gp),
                                               _ptr_array(gl), _lib.stream_ptr()), "tir_vm_app_products_bwd")
        return (None, None, *_to_param_layout(gp, gl))


def app_products(model, xn):
    """[n, 3*C] plane*line products of the appearance factors (tensoRF_rotated_lights.py:141-153)."""
    xn = xn.detach().reshape(-1, 3).float().contiguous()
    params = list(model.app_plane) + list(model.app_line)
    if _needs_grad(params):
        return _AppProducts.apply(model, xn, *params)
    with torch.no_grad():
        return _AppProducts.apply(model, xn, *params)


class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sigma, dist, offsets, scale, z):
        lib = _lib.load()
        sigma = sigma.contiguous()
        n_rays = offsets.numel() - 1
        weight = torch.zeros_like(sigma)      # rows of a static-capacity list past the real total stay 0
        trans = torch.zeros_like(sigma)
        t_last = torch.ones(n_rays, device=sigma.device)
        acc = torch.zeros(n_rays, device=sigma.device)
        depth = torch.zeros(n_rays, device=sigma.device)
        if sigma.numel() > 0:        # a chunk whose rays all miss the aabb has no valid sample at all
            _lib.check(lib.tir_composite_fwd(_lib.dptr(sigma), _lib.dptr(dist), _lib.dptr(offsets, torch.int64),
                                             n_rays, float(scale), _lib.dptr(weight), _lib.dptr(trans),
                                             _lib.dptr(t_last), sigma.numel(), _lib.dptr(z), _lib.dptr(acc),
                                             _lib.dptr(depth), _lib.stream_ptr()), "tir_composite_fwd")
        ctx.save_for_backward(sigma, dist, offsets, weight, trans, z)
        ctx.scale = float(scale)
        ctx.mark_non_differentiable(t_last)
        return weight, t_last, acc, depth

    @staticmethod
    def backward(ctx, g_weight, _g_t, g_acc, g_depth):
        sigma, dist, offsets, weight, trans, z = ctx.saved_tensors
        lib = _lib.load()
        nul = C.c_void_p(0)

        def ptr(t):
            return nul if t is None else _lib.dptr(t.contiguous().float())
        g_sigma = torch.zeros_like(sigma)
        if sigma.numel() > 0:
            _lib.check(lib.tir_composite_bwd(_lib.dptr(sigma), _lib.dptr(dist), _lib.dptr(offsets, torch.int64),
                                             offsets.numel() - 1, ctx.scale, _lib.dptr(weight), _lib.dptr(trans),
                                             ptr(g_weight), _lib.dptr(g_sigma), sigma.numel(), _lib.dptr(z),
                                             ptr(g_acc), ptr(g_depth), _lib.stream_ptr()), "tir_composite_bwd")
        return g_sigma, None, None, None, None


def composite(sigma, dist, offsets, scale, z):
    """raw2alpha over ray segments (tensorBase:21-28) + the two per-ray sums that need every valid sample
    (acc_map, depth_map, tensorBase:974-975) -> (weight [n_valid], T_last [n_rays], acc [n_rays], depth [n_rays])."""
    return _Composite.apply(sigma, dist, offsets, scale, z)


def valid_samples(model, rays_o, rays_d, *, n_samples=-1, jitter=None, table=None, counters=None, no_bbox=False,
                  count_only=False, capacity=None):
    """Ray-sorted list of the samples that are inside the aabb and pass the alpha mask.
    -> dict(ray, sample, xn, z, dist, offsets, counts).
    ``capacity``: static list length (no host sync; rows past the real total are zero padding, rows that do not fit
    are dropped and flagged in ``overflow``) — the shape-static form used under CUDA-graph capture."""
    lib = _lib.load()
    f = ops.device_field(model).refresh(model)
    keep = []
    cfg = ops.march_cfg(model, table=table, n_samples=n_samples, jitter=jitter, keep=keep)
    if no_bbox:
        cfg.flags = _lib.MARCH_NO_BBOX
    ro = rays_o.detach().reshape(-1, 3).float().contiguous()
    rd = rays_d.detach().reshape(-1, 3).float().contiguous()
    n = ro.shape[0]
    dev = ro.device
    counts = torch.zeros(n, dtype=torch.int32, device=dev)
    _lib.check(lib.tir_valid_samples_count(C.byref(f), _lib.dptr(ro), _lib.dptr(rd), n, C.byref(cfg),
                                           _lib.dptr(counts, torch.int32),
                                           None if counters is None else _lib.dptr(counters, torch.int64),
                                           _lib.stream_ptr()), "tir_valid_samples_count")
    if count_only:
        return {"counts": counts}
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=offsets[1:])
    if capacity is None:
        total = int(offsets[-1].item())
        alloc = torch.empty
        cap_arg = 0
    else:
        total = int(capacity)
        alloc = torch.zeros
        cap_arg = total
    out = {"counts": counts, "offsets": offsets,
           "ray": alloc(total, dtype=torch.int32, device=dev),
           "sample": alloc(total, dtype=torch.int32, device=dev),
           "xn": alloc((total, 3), device=dev), "z": alloc(total, device=dev),
           "dist": alloc(total, device=dev)}
    if capacity is not None:
        out["overflow"] = offsets[-1] > total
    if total > 0:
        _lib.check(lib.tir_valid_samples_fill(C.byref(f), _lib.dptr(ro), _lib.dptr(rd), n, C.byref(cfg),
                                              _lib.dptr(offsets, torch.int64), _lib.dptr(out["ray"], torch.int32),
                                              _lib.dptr(out["sample"], torch.int32), _lib.dptr(out["xn"]),
                                              _lib.dptr(out["z"]), _lib.dptr(out["dist"]), cap_arg, _lib.stream_ptr()),
                   "tir_valid_samples_fill")
    return out
