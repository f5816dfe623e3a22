"""Kernel-level timing of the hot path on one GPU (development tool; bench.py is the contract)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_scene(G, dev, mask_res=None):
    from tensoir_b200 import TensorVMSplit
    from tensoir_b200.synthetic import install_lego_density
    torch.manual_seed(20211202)
    aabb = torch.tensor([[-1.5] * 3, [1.5] * 3], device=dev)
    m = TensorVMSplit(aabb, [G] * 3, dev, density_n_comp=[16] * 3, appearance_n_comp=[48] * 3, app_dim=27,
                      shadingMode='MLP_Fea', step_ratio=0.5, normals_kind='derived_plus_predicted',
                      light_rotation=['000'], light_kind='sg', alphaMask_thres=0.001)
    install_lego_density(m)
    r = mask_res or min(G, 256)
    m.updateAlphaMask((r, r, r))
    return m


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    dev = "cuda:0"
    import __graft_entry__ as g
    g.build()
    from tensoir_b200 import ops, Renderer_TensoIR_train
    from tensoir_b200.synthetic import hemisphere_poses, training_batch
    from gpu_helpers import renderer_args
    m = build_scene(a.grid, dev)
    rays, li = training_batch(hemisphere_poses(100), a.batch, 0)
    n_s = min(1000000, int((3 * a.grid ** 2) ** 0.5 / 0.5))
    out = {"grid": a.grid, "batch": a.batch, "n_samples": n_s}
    with torch.no_grad():
        ret = Renderer_TensoIR_train(rays, None, li, m, N_samples=n_s, white_bg=True, is_train=False, is_relight=True,
                                     sample_method='fixed_envirmap', device=dev, args=renderer_args(96))
    mask = ret["acc_map"] > 0.5
    out["surface_hits"] = int(mask.sum())
    rd = rays.to(dev)
    surf = (rd[:, :3] + ret["depth_map"][:, None] * rd[:, 3:])[mask]
    nrm = ret["normal_map"][mask]
    lid = li.to(dev)[mask]
    dirs = m.gen_light_incident_dirs(method='fixed_envirmap').to(dev)
    cnt = ops.new_counters(dev)
    ops.secondary_radiance(m, surf, nrm, lid, dirs, counters=cnt)
    c = ops.counters_dict(cnt)
    out["secondary_counters"] = c
    t_sec = timed(lambda: ops.secondary_radiance(m, surf, nrm, lid, dirs), a.iters)
    bytes_alg = 32 * c["mask"] + 1152 * c["density"] + 3456 * c["app"] + 16 * c["rays"]
    out["secondary_ms"] = t_sec
    out["secondary_rays_per_s"] = c["rays"] / (t_sec * 1e-3)
    out["secondary_alg_GBps"] = bytes_alg / (t_sec * 1e-3) / 1e9
    # split: march only (density) vs march+list vs mlp
    table = ops.equal_z_table(96, 0.05, 1.5, dev)
    P = surf[:, None, :].expand(-1, dirs.shape[0], -1).reshape(-1, 3).contiguous()
    D = dirs[None].expand(surf.shape[0], -1, -1).reshape(-1, 3).contiguous()
    out["march_density_all_dirs_ms"] = timed(lambda: ops.march_density(m, P, D, table=table), a.iters)
    # primary
    cnt.zero_()
    m.__dict__["_tir_counters"] = cnt
    with torch.no_grad():
        m(rd, li.to(dev), is_train=False, is_relight=True, N_samples=n_s)
    out["primary_counters"] = ops.counters_dict(cnt)
    del m.__dict__["_tir_counters"]

    def prim():
        with torch.no_grad():
            m(rd, li.to(dev), is_train=False, is_relight=True, N_samples=n_s)
    out["primary_eval_ms"] = timed(prim, a.iters)
    out["march_density_primary_ms"] = timed(lambda: ops.march_density(m, rd[:, :3], rd[:, 3:], n_samples=n_s), a.iters)

    def full_eval():
        with torch.no_grad():
            Renderer_TensoIR_train(rays, None, li, m, N_samples=n_s, white_bg=True, is_train=False, is_relight=True,
                                   sample_method='fixed_envirmap', device=dev, args=renderer_args(96))
    out["renderer_eval_ms"] = timed(full_eval, a.iters)

    opt = torch.optim.Adam(m.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))

    def train_step():
        ret = Renderer_TensoIR_train(rays, None, li, m, N_samples=n_s, white_bg=True, is_train=True, is_relight=True,
                                     sample_method='stratified_sampling', device=dev, args=renderer_args(96))
        loss = ((ret["rgb_map"] - 0.5) ** 2).mean() + 0.2 * ((ret["rgb_with_brdf_map"] - 0.5) ** 2).mean() \
            + 0.0005 * ret["normals_diff_map"].mean() + 0.001 * ret["normals_orientation_loss_map"].mean() \
            + 0.001 * ret["albedo_smoothness_loss"] + 0.001 * ret["roughness_smoothness_loss"]
        opt.zero_grad()
        loss.backward()
        opt.step()
    out["train_step_ms"] = timed(train_step, a.iters)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
