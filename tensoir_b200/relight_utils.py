"""Host-side mirror of models/relight_utils.py: the physically-based shading integral.

Secondary visibility / indirect-light rays — ~99 % of the work of a relight step —
run in the fused CUDA march (csrc/tir_march.cu + csrc/tir_mlp.cu), generated on chip from
(surface point, direction) pairs; the small differentiable epilogue (GGX, SG lights, quadrature, sRGB)
stays in PyTorch so autograd reaches normal / albedo / roughness / lgtSGs exactly as in the reference
(the secondary march is @torch.no_grad in the reference too, relight_utils.py:344, :777).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from torch.profiler import record_function

from . import ops


def safe_l2_normalize(x, dim=None, eps=1e-6):
    return F.normalize(x, p=2, dim=dim, eps=eps)


def GGX_specular(normal, pts2c, pts2l, roughness, fresnel):
    """relight_utils.py:17-50."""
    L = F.normalize(pts2l, dim=-1)
    V = F.normalize(pts2c, dim=-1)
    H = F.normalize((L + V[:, None, :]) / 2.0, dim=-1)
    N = F.normalize(normal, dim=-1)
    NoV = torch.sum(V * N, dim=-1, keepdim=True)
    N = N * NoV.sign()
    NoL = torch.sum(N[:, None, :] * L, dim=-1, keepdim=True).clamp_(1e-6, 1)
    NoV = torch.sum(N * V, dim=-1, keepdim=True).clamp_(1e-6, 1)
    NoH = torch.sum(N[:, None, :] * H, dim=-1, keepdim=True).clamp_(1e-6, 1)
    VoH = torch.sum(V[:, None, :] * H, dim=-1, keepdim=True).clamp_(1e-6, 1)
    alpha = roughness * roughness
    alpha2 = alpha * alpha
    k = (alpha + 2 * roughness + 1.0) / 8.0
    FMi = ((-5.55473) * VoH - 6.98316) * VoH
    frac0 = fresnel[:, None, :] + (1 - fresnel[:, None, :]) * torch.pow(2.0, FMi)
    frac = frac0 * alpha2[:, None, :]
    nom0 = NoH * NoH * (alpha2[:, None, :] - 1) + 1
    nom1 = NoV * (1 - k) + k
    nom2 = NoL * (1 - k[:, None, :]) + k[:, None, :]
    nom = (4 * np.pi * nom0 * nom0 * nom1[:, None, :] * nom2).clamp_(1e-6, 4 * np.pi)
    return frac / nom


brdf_specular = GGX_specular


def linear2srgb_torch(tensor_0to1):
    """relight_utils.py:489-515; the clip of :518-533 is applied unconditionally (clamp is the identity on
    in-range data), which removes the reference's min/max host sync."""
    t = torch.clamp(tensor_0to1, min=0, max=1)
    lin = t * 12.92
    nonlin = 1.055 * torch.pow(t + 1e-6, 1 / 2.4) - (1.055 - 1)
    return torch.where(t <= 0.0031308, lin, nonlin)


@torch.no_grad()
def compute_transmittance(tensoIR, surf_pts, light_in_dir, nSample=128, vis_near=0.1, vis_far=2, device='cuda'):
    """relight_utils.py:657-705 -> (nerv_vis, nerfactor_vis): one fused density march."""
    table = ops.equal_z_table(nSample, vis_near, vis_far, surf_pts.device)
    t_last, acc, _ = ops.march_density(tensoIR, surf_pts, light_in_dir, table=table,
                                       counters=tensoIR.__dict__.get("_tir_counters"))
    return t_last, 1 - acc


@torch.no_grad()
def compute_radiance(tensoIR, surf_pts, light_in_dir, light_idx, nSample=128, vis_near=0.05, vis_far=1.5,
                     device=None):
    """relight_utils.py:777-834 -> (nerv_vis, nerfactor_vis, indirect_light)."""
    table = ops.equal_z_table(nSample, vis_near, vis_far, surf_pts.device)
    t_last, acc, _, rgb, _ = ops.march_radiance(tensoIR, surf_pts, light_in_dir, light_idx, table=table,
                                                counters=tensoIR.__dict__.get("_tir_counters"))
    return t_last, 1 - acc, rgb


@torch.no_grad()
def compute_secondary_shading_effects(tensoIR, surface_pts, surf2light, light_idx, nSample=96, vis_near=0.05,
                                      vis_far=1.5, chunk_size=15000, device='cuda'):
    """relight_utils.py:344-399 over explicit (point, direction) rays.  ``chunk_size`` is accepted and
    ignored: the fused march never materialises [rays, samples, 3]."""
    vis, _, ind = compute_radiance(tensoIR, surface_pts, surf2light, light_idx, nSample, vis_near, vis_far)
    return vis.reshape(-1, 1), ind.reshape(-1, 3)


def render_with_BRDF(depth_map, normal_map, albedo_map, roughness_map, fresnel_map, rays, tensoIR, light_idx,
                     sample_method='fixed_envirmap', chunk_size=15000, device='cuda', use_linear2srgb=True,
                     args=None):
    """relight_utils.py:403-483."""
    device = depth_map.device
    rays_o, rays_d = rays[..., :3].to(device), rays[..., 3:].to(device)
    surface_xyz = rays_o + depth_map.unsqueeze(-1) * rays_d
    light_area_weight = tensoIR.__dict__.get("_tir_area_weight_dev")
    if light_area_weight is None or light_area_weight.device != device:
        light_area_weight = tensoIR.light_area_weight.to(device)
        tensoIR.__dict__["_tir_area_weight_dev"] = light_area_weight
    st = tensoIR.__dict__.get("_tir_static")
    if st is not None:
        incident_light_dirs = st["dirs"]         # device buffer refilled by the host (same generator order)
    else:
        incident_light_dirs = tensoIR.gen_light_incident_dirs(method=sample_method).to(device)   # host draws, ref order
    nlights = incident_light_dirs.shape[0]
    surf2c = safe_l2_normalize(-rays_d, dim=-1)
    # secondary rays: cosine test + 96-sample march + appearance MLP, generated on chip per (point, direction)
    with record_function("tir::secondary"):
        vis, indirect, _ = ops.secondary_radiance(
            tensoIR, surface_xyz, normal_map, light_idx, incident_light_dirs,
            n_sample=args.second_nSample, near=args.second_near, far=args.second_far,
            counters=tensoIR.__dict__.get("_tir_counters"))
    with record_function("tir::shade_epilogue"):
        return _shade(tensoIR, normal_map, albedo_map, roughness_map, fresnel_map, surf2c, vis, indirect,
                      incident_light_dirs, light_idx, light_area_weight, nlights, sample_method, use_linear2srgb, device)


def render_hits(depth_map, normal_map, albedo_map, roughness_map, fresnel_map, acc_mask, rays, tensoIR, light_idx,
                sample_method='fixed_envirmap', use_linear2srgb=True, args=None):
    """``rgb_with_brdf`` of Renderer_TensoIR_train (renderer.py:99-106) for the WHOLE ray batch at once:
        out = ones;  out[acc_mask] = render_with_BRDF(depth_map[acc_mask], ...)
    without ever compacting the hits: non-hit rays get a zero normal (the secondary march's cosine test then skips all
    their directions) and shade to the constant 1 inside the fused kernel (csrc/tir_shade.cu, tir_shade_hits_*), which
    also applies the [0,1] clip and linear2srgb.  Nothing depends on the number of hits, so there is no list to size
    (or overflow) under CUDA-graph capture, and no boolean-mask gathers / scatters in either direction."""
    import ctypes as C
    from . import _lib
    from .shade import shade_hits
    device = depth_map.device
    lib = _lib.load()
    rays = rays.float().contiguous()
    n = rays.shape[0]
    st = tensoIR.__dict__.get("_tir_static")
    if st is not None:
        dirs = st["dirs"]                                   # device buffer refilled by the host (same generator order)
    else:
        dirs = tensoIR.gen_light_incident_dirs(method=sample_method).to(device)     # host draws, reference order
    weight = tensoIR.__dict__.get("_tir_area_weight_dev")
    if weight is None or weight.device != device:
        weight = tensoIR.light_area_weight.to(device)
        tensoIR.__dict__["_tir_area_weight_dev"] = weight
    nl = dirs.shape[0]
    if sample_method == 'stratifed_sample_equal_areas':
        weight = torch.full((nl,), 4 * torch.pi / nl, device=device)
    surf = torch.empty(n, 3, device=device)
    nrm = torch.empty(n, 3, device=device)
    with torch.no_grad():
        nm = normal_map.detach().float().contiguous()
        _lib.check(lib.tir_hits_prepare(_lib.dptr(rays), _lib.dptr(depth_map.detach().float().contiguous()),
                                        _lib.dptr(nm), _lib.dptr(acc_mask.contiguous(), torch.bool), n, _lib.dptr(surf),
                                        _lib.dptr(nrm), _lib.stream_ptr()), "tir_hits_prepare")
        with record_function("tir::secondary"):
            vis, indirect, _ = ops.secondary_radiance(
                tensoIR, surf, nrm, light_idx, dirs, n_sample=args.second_nSample, near=args.second_near,
                far=args.second_far, counters=tensoIR.__dict__.get("_tir_counters"))
    direct = tensoIR.get_light_rgbs(dirs, device=device).to(device)            # [L, nl, 3], autograd reaches lgtSGs
    with record_function("tir::shade_epilogue"):
        return shade_hits(normal_map, albedo_map, roughness_map, fresnel_map, direct, rays, acc_mask, light_idx, dirs,
                          weight, vis, indirect, srgb=use_linear2srgb)


def _shade(tensoIR, normal_map, albedo_map, roughness_map, fresnel_map, surf2c, vis, indirect, incident_light_dirs,
           light_idx, light_area_weight, nlights, sample_method, use_linear2srgb, device):
    """Quadrature of the rendering equation, relight_utils.py:452-483: fused CUDA kernel (forward + analytic
    backward, csrc/tir_shade.cu); the SG light table stays in PyTorch (autograd reaches lgtSGs through it)."""
    from .shade import shade
    envir_map_light_rgbs = tensoIR.get_light_rgbs(incident_light_dirs, device=device).to(device)   # [L, nl, 3]
    if sample_method == 'stratifed_sample_equal_areas':
        weight = torch.full((nlights,), 4 * torch.pi / nlights, device=device)     # mean(4 pi x) over directions
    else:
        weight = light_area_weight
    rgb_with_brdf = shade(normal_map, albedo_map, roughness_map, fresnel_map, envir_map_light_rgbs, surf2c, light_idx,
                          incident_light_dirs, weight, vis, indirect)
    rgb_with_brdf = torch.clamp(rgb_with_brdf, min=0.0, max=1.0)
    if use_linear2srgb and rgb_with_brdf.shape[0] > 0:
        rgb_with_brdf = linear2srgb_torch(rgb_with_brdf)
    return rgb_with_brdf


# ---------------------------------------------------------------------------------------------------------------------
# Rest of the module surface the reference's data loaders / scripts import from models.relight_utils
# (dataLoader/tensoIR_rotation_setting.py:13, scripts/relight_importance.py:21-24)
# ---------------------------------------------------------------------------------------------------------------------
def grid_sample(image, optical):
    """Differentiable bilinear sampler with CLAMPED tap indices and unclamped weights (relight_utils.py:57-107): the
    sampler compute_densityfeature_with_xyz_grad differentiates twice.  image [N,C,IH,IW], optical [N,H,W,2] in
    [-1,1] (align_corners=True convention) -> [N,C,H,W].  The hot path evaluates this analytically in
    csrc/tir_vm.cu (tir_vm_density_grad); this tensor version serves callers that use it directly."""
    N, C, IH, IW = image.shape
    _, H, W, _ = optical.shape
    x = (optical[..., 0] + 1) * 0.5 * (IW - 1)
    y = (optical[..., 1] + 1) * 0.5 * (IH - 1)
    with torch.no_grad():
        x0, y0 = torch.floor(x), torch.floor(y)
    fx, fy = x - x0, y - y0                               # weights come from the UNclamped corners
    flat = image.reshape(N, C, IH * IW)

    def tap(dx, dy):
        with torch.no_grad():
            xi = (x0 + dx).clamp(0, IW - 1)
            yi = (y0 + dy).clamp(0, IH - 1)
            idx = (yi * IW + xi).long().view(N, 1, H * W).expand(N, C, H * W)
        return torch.gather(flat, 2, idx).view(N, C, H, W)
    w00, w10 = ((1 - fx) * (1 - fy)).view(N, 1, H, W), (fx * (1 - fy)).view(N, 1, H, W)
    w01, w11 = ((1 - fx) * fy).view(N, 1, H, W), (fx * fy).view(N, 1, H, W)
    return tap(0, 0) * w00 + tap(1, 0) * w10 + tap(0, 1) * w01 + tap(1, 1) * w11


def _clip_0to1_warn_torch(tensor_0to1):
    """relight_utils.py:518-533 (the warning is dropped: clamp is the identity on in-range data, no host sync)."""
    if isinstance(tensor_0to1, torch.Tensor):
        return torch.clamp(tensor_0to1, min=0, max=1)
    if isinstance(tensor_0to1, np.ndarray):
        return np.clip(tensor_0to1, 0, 1)
    raise NotImplementedError(f'Do not support dtype {type(tensor_0to1)}')


def _convert_sph_conventions(pts_r_angle1_angle2, what2what):
    """relight_utils.py:537-567: (r, lat, lng) <-> (r, theta, phi)."""
    p = np.asarray(pts_r_angle1_angle2)
    out = np.zeros(p.shape)
    out[:, 0] = p[:, 0]
    out[:, 1] = np.pi / 2 - p[:, 1]
    if what2what == 'lat-lng_to_theta-phi':
        out[:, 2] = np.where(p[:, 2] < 0, 2 * np.pi + p[:, 2], p[:, 2])
    elif what2what == 'theta-phi_to_lat-lng':
        out[:, 2] = np.where(p[:, 2] > np.pi, p[:, 2] - 2 * np.pi, p[:, 2])
    else:
        raise NotImplementedError(what2what)
    return out


def sph2cart(pts_sph, convention='lat-lng'):
    """relight_utils.py:570-595."""
    pts_sph = np.asarray(pts_sph)
    assert pts_sph.ndim == 2 and pts_sph.shape[-1] == 3, "Shape of input mush be (n, 3)"
    assert (np.abs(pts_sph[:, 1:]) <= 2 * np.pi).all(), "Input degree falls out of [-2pi, 2pi]"
    if convention == 'theta-phi':
        pts_sph = _convert_sph_conventions(pts_sph, 'theta-phi_to_lat-lng')
    elif convention != 'lat-lng':
        raise NotImplementedError(convention)
    r, lat, lng = pts_sph[:, 0], pts_sph[:, 1], pts_sph[:, 2]
    return np.stack((r * np.cos(lat) * np.cos(lng), r * np.cos(lat) * np.sin(lng), r * np.sin(lat)), axis=-1)


@torch.no_grad()
def sample_ray_equally(tensoIR, rays_o, rays_d, nSample=-1, vis_near=0.03, vis_far=1.5, device=None):
    """relight_utils.py:707-722 as tensors (the fused marches generate these samples on chip; this form is for callers
    that want the points themselves) -> (rays_pts [N,S,3], z_vals [1,S], in-aabb mask [N,S])."""
    z_vals = ops.equal_z_table(nSample, vis_near, vis_far, rays_o.device).unsqueeze(0)
    rays_pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals.view(1, -1, 1)
    aabb = tensoIR.aabb.to(rays_o.device)
    outside = ((aabb[0] > rays_pts) | (rays_pts > aabb[1])).any(dim=-1)
    return rays_pts, z_vals, ~outside


def _pairs(pts, light_xyz):
    surf2light = safe_l2_normalize(light_xyz[None, :, :] - pts[:, None, :], dim=-1)
    return pts.unsqueeze(1).expand(-1, light_xyz.shape[0], -1).reshape(-1, 3), surf2light.reshape(-1, 3)


@torch.no_grad()
def compute_visibility(tensoIR, pts, light_xyz, nSample, vis_near, vis_far, args, device='cuda'):
    """relight_utils.py:617-655: visibility of every (surface point, light position) pair -> [N, n_lights, 1].
    One fused density march over all pairs (no 81920-ray chunk loop: nothing [rays, samples]-shaped exists)."""
    surface_pts, surf2light = _pairs(pts, light_xyz)
    nerv_vis, nerfactor_vis = compute_transmittance(tensoIR, surface_pts, surf2light, nSample, vis_near, vis_far)
    vis = nerv_vis if args.vis_equation == 'nerv' else nerfactor_vis
    return vis.reshape(-1, light_xyz.shape[0], 1)


@torch.no_grad()
def compute_visibility_and_indirect_light(tensoIR, pts, light_xyz, light_idx, nSample, vis_near, vis_far, args,
                                          device='cuda'):
    """relight_utils.py:727-775 -> (visibility [N, n_lights, 1], indirect_light [N, n_lights, 3])."""
    surface_pts, surf2light = _pairs(pts, light_xyz)
    li = light_idx.view(-1, 1, 1).expand((-1, light_xyz.shape[0], 1)).reshape(-1, 1)
    nerv_vis, nerfactor_vis, indirect = compute_radiance(tensoIR, surface_pts, surf2light, li, nSample, vis_near,
                                                         vis_far)
    vis = nerv_vis if args.vis_equation == 'nerv' else nerfactor_vis
    return vis.reshape(-1, light_xyz.shape[0], 1), indirect.reshape(-1, light_xyz.shape[0], 3)


def __getattr__(name):
    # read_hdr / Environment_Light live next to the relighting driver (tensoir_b200/relight.py, which imports this
    # module); resolved lazily to keep the import graph acyclic
    if name in ("read_hdr", "Environment_Light"):
        from . import relight
        return getattr(relight, name)
    raise AttributeError(name)
