"""The drop-in boundary: renderer.Renderer_TensoIR_train (renderer.py:57-127) and the TensoRF-heritage
OctreeRender_trilinear_fast alias named by BASELINE.json (absent from this reference)."""
from __future__ import annotations

import torch

from .relight_utils import render_with_BRDF


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
    if is_relight and tensoIR.__dict__.get("_tir_static") is not None:
        # shape-static variant (CUDA-graph capture): the acc_mask rows are compacted into a list of STATIC capacity
        # (padding rows point at a dummy row N, carry a zero normal so the secondary kernel's cosine test skips all
        # their directions, and are dropped by the scatter), exactly mirroring the dynamic branch below.
        st = tensoIR.__dict__["_tir_static"]
        n = rgb_map.shape[0]
        cap = min(int(st.get("cap_hit", n)), n) if n > 0 else 0
        idx = torch.nonzero_static(acc_mask, size=cap, fill_value=-1).reshape(-1)
        real = idx >= 0
        from .primary import note_count
        note_count(st, 2, acc_mask.sum(), cap)
        src = idx.clamp(min=0)
        keep = real[:, None].to(normal_map.dtype)
        shaded = render_with_BRDF(depth_map.index_select(0, src), normal_map.index_select(0, src) * keep,
                                  albedo_map.index_select(0, src), roughness_map.index_select(0, src).repeat(1, 3),
                                  fresnel_map.index_select(0, src), rays.index_select(0, src), tensoIR,
                                  light_idx.index_select(0, src), sample_method, chunk_size=chunk_size, device=device,
                                  args=args)
        dst = torch.where(real, idx, torch.full_like(idx, n))
        rgb_with_brdf = torch.ones((n + 1, 3), device=rgb_map.device, dtype=rgb_map.dtype).index_copy(0, dst, shaded)[:n]
    elif is_relight:
        masked = render_with_BRDF(depth_map[acc_mask], normal_map[acc_mask], albedo_map[acc_mask],
                                  roughness_map[acc_mask].repeat(1, 3), fresnel_map[acc_mask], rays[acc_mask],
                                  tensoIR, light_idx[acc_mask], sample_method, chunk_size=chunk_size,
                                  device=device, args=args)
        rgb_with_brdf = torch.ones_like(rgb_map)
        rgb_with_brdf[acc_mask] = masked
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
