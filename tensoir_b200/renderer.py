"""The drop-in boundary: renderer.Renderer_TensoIR_train (renderer.py:57-127) and the TensoRF-heritage
OctreeRender_trilinear_fast alias named by BASELINE.json (absent from this reference)."""
from __future__ import annotations

import torch

from .relight_utils import render_hits, render_with_BRDF  # noqa: F401  (renderer.py:6 star-imports it too)


def Renderer_TensoIR_train(rays=None, normal_gt=None, light_idx=None, tensoIR=None, N_samples=-1, ndc_ray=False,
                           white_bg=True, is_train=False, is_relight=True, sample_method='fixed_envirmap',
                           chunk_size=15000, device='cuda', args=None):
    """Same signature, same 12-entry dict as renderer.py:57-127.  Host buffers in -> H2D of rays / light_idx
    here (renderer.py:74-75), everything after that stays on the device."""
    rays = rays.to(device, non_blocking=True)
    light_idx = light_idx.to(device, torch.int32, non_blocking=True)
    (rgb_map, depth_map, normal_map, albedo_map, roughness_map, fresnel_map, acc_map, normals_diff_map,
     normals_orientation_loss_map, acc_mask, albedo_smoothness_loss, roughness_smoothness_loss) = \
        tensoIR(rays, light_idx, is_train=is_train, white_bg=white_bg, is_relight=is_relight, ndc_ray=ndc_ray,
                N_samples=N_samples)
    if tensoIR.normals_kind == "gt_normals" and normal_gt is not None:
        normal_map = normal_gt.to(device)
    if is_relight and rgb_map.shape[0] > 0:
        # rgb_with_brdf = ones; rgb_with_brdf[acc_mask] = render_with_BRDF(<maps>[acc_mask], ...) (renderer.py:99-106)
        # evaluated for the whole batch at once, non-hit rays shading to 1 inside the kernel: no list of hits exists
        rgb_with_brdf = render_hits(depth_map, normal_map, albedo_map, roughness_map, fresnel_map, acc_mask, rays,
                                    tensoIR, light_idx, sample_method, args=args)
    else:
        rgb_with_brdf = torch.ones_like(rgb_map)
    return {"rgb_map": rgb_map, "depth_map": depth_map, "normal_map": normal_map, "albedo_map": albedo_map,
            "acc_map": acc_map, "roughness_map": roughness_map, "fresnel_map": fresnel_map,
            "rgb_with_brdf_map": rgb_with_brdf, "normals_diff_map": normals_diff_map,
            "normals_orientation_loss_map": normals_orientation_loss_map,
            "albedo_smoothness_loss": albedo_smoothness_loss, "roughness_smoothness_loss": roughness_smoothness_loss}


def OctreeRender_trilinear_fast(rays, tensorf, chunk=4096, N_samples=-1, ndc_ray=False, white_bg=True,
                                is_train=False, device='cuda'):
    """TensoRF's renderer contract for the density/appearance-only model (models/tensoRF_init.TensorVMSplit,
    forward at tensorBase_init.py:406-462): chunked loop -> (rgbs, None, depth_maps, None, None)."""
    rgbs, depth_maps = [], []
    n = rays.shape[0]
    for i in range(n // chunk + int(n % chunk > 0)):
        rays_chunk = rays[i * chunk:(i + 1) * chunk].to(device)
        rgb_map, depth_map = tensorf(rays_chunk, is_train=is_train, white_bg=white_bg, ndc_ray=ndc_ray,
                                     N_samples=N_samples)
        rgbs.append(rgb_map)
        depth_maps.append(depth_map)
    if not rgbs:                      # zero rays: empty maps instead of torch.cat([]) raising
        return (torch.zeros((0, 3), device=device), None, torch.zeros((0,), device=device), None, None)
    return torch.cat(rgbs), None, torch.cat(depth_maps), None, None
