"""Shared test helpers: build oracle fields from golden fixtures."""
import torch

from oracle import tensoir_oracle as O


def oracle_field(fx, with_mask=True, requires_grad=False):
    kw = {}
    if fx["kind"] == "rotated":
        kw["light_rotation"] = fx["light_rotation"]
    if fx["kind"] == "general":
        kw["lgt_sgs_list"] = fx["lgt_sgs_list"]
    f = O.field_from_state_dict(fx["state_dict"], fx["aabb"], fx["grid_size"], kind=fx["kind"],
                                alpha_volume=fx["alpha_volume"] if with_mask else None,
                                alpha_aabb=fx["alpha_aabb"] if with_mask else None, **kw)
    if requires_grad:
        for p in f.parameters():
            p.requires_grad_(True)
    return f


def named_oracle_params(f):
    """reference state_dict key -> oracle tensor."""
    out = {}
    for i in range(3):
        out[f"density_plane.{i}"] = f.density_plane[i]
        out[f"density_line.{i}"] = f.density_line[i]
        out[f"app_plane.{i}"] = f.app_plane[i]
        out[f"app_line.{i}"] = f.app_line[i]
    out["basis_mat.weight"] = f.basis_w
    if f.light_line_w is not None:
        out["light_line.weight"] = f.light_line_w
    for name, m in (("renderModule", f.mlp_rgb), ("renderModule_brdf", f.mlp_brdf),
                    ("renderModule_normal", f.mlp_normal)):
        if m is None:
            continue
        for j, (w, b) in zip((0, 2, 4), ((m.w0, m.b0), (m.w1, m.b1), (m.w2, m.b2))):
            out[f"{name}.mlp.{j}.weight"] = w
            out[f"{name}.mlp.{j}.bias"] = b
    if f.lgt_sgs is not None and f.light_rotation is not None:
        out["lgtSGs"] = f.lgt_sgs[0]
    return out
