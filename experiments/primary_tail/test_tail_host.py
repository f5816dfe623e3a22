"""CPU check of experiments/primary_tail/tail_body.h: the host build against torch autograd of the expressions the
product currently evaluates with ~90 small kernels (tensoir_b200/primary.py:_forward_relight_tail).

    python -m pytest experiments/primary_tail/test_tail_host.py -q
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("tail") / "libtail_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "tail_host.cpp")],
                   check=True)
    return C.CDLL(so)


def rel(a, b):          # TensorBase.compute_relative_smoothness_loss
    base = torch.maximum(a, b).clip(min=1e-6)
    return torch.sum(((a - b) / base) ** 2, dim=-1, keepdim=True)


def torch_tail(w, ray, rgb, brdf, brdfj, vn, dn, viewdirs, n_rays):
    v_alb, v_rough = brdf[..., :3], brdf[..., 3:4] * 0.9 + 0.09
    a_cost = rel(v_alb, brdfj[..., :3])
    r_cost = rel(v_rough, brdfj[..., 3:4] * 0.9 + 0.09)
    vd = viewdirs.index_select(0, ray)
    if dn is not None:
        nd = torch.sum(torch.pow(vn - dn, 2), dim=-1, keepdim=True)
        no = torch.sum(vd * vn, dim=-1, keepdim=True).clamp(min=0)
    else:
        nd = no = torch.zeros_like(a_cost)
    vals = torch.cat([rgb, vn, v_alb, v_rough, a_cost, r_cost, nd, no], dim=-1)
    return torch.zeros(n_rays, 14).index_add_(0, ray, w[:, None] * vals)


def p(t, ct=C.c_float):
    return None if t is None else t.ctypes.data_as(C.POINTER(ct))


@pytest.mark.parametrize("both", [True, False])
def test_tail_matches_autograd(lib, both):
    torch.manual_seed(3)
    n, n_rays = 500, 40
    ray = torch.sort(torch.randint(0, n_rays, (n,)))[0]
    w = torch.rand(n) * 0.3
    w[::17] = 0.0                                                     # padding rows of the static-capacity list
    rgb, brdf, brdfj = torch.rand(n, 3), torch.rand(n, 4), torch.rand(n, 4)
    brdfj[5] = brdf[5]                                                # ties in maximum()
    vn = torch.nn.functional.normalize(torch.randn(n, 3), dim=-1)
    dn = torch.nn.functional.normalize(torch.randn(n, 3), dim=-1) if both else None
    viewdirs = torch.nn.functional.normalize(torch.randn(n_rays, 3), dim=-1)
    leaves = [t.clone().requires_grad_(True) for t in (w, rgb, brdf, brdfj, vn)] + \
             ([dn.clone().requires_grad_(True)] if both else [])
    want = torch_tail(leaves[0], ray, leaves[1], leaves[2], leaves[3], leaves[4], leaves[5] if both else None,
                      viewdirs, n_rays)
    G = torch.randn(n_rays, 14)
    want.backward(G)

    arr = [t.numpy().astype(np.float32) for t in (w, rgb, brdf, brdfj, vn)]
    dn_a = dn.numpy().astype(np.float32) if both else None
    ray_a, vd_a = ray.numpy().astype(np.int64), viewdirs.numpy().astype(np.float32)
    packed = np.zeros((n_rays, 14), np.float32)
    lib.tail_forward_host(C.c_int64(n), p(arr[0]), p(ray_a, C.c_int64), p(arr[1]), p(arr[2]), p(arr[3]), p(arr[4]),
                          p(dn_a), p(vd_a), p(packed))
    assert np.allclose(packed, want.detach().numpy(), rtol=2e-5, atol=2e-6)

    G_a = G.numpy().astype(np.float32)
    g = [np.zeros_like(a) for a in arr]
    g_dn = np.zeros((n, 3), np.float32) if both else None
    lib.tail_backward_host(C.c_int64(n), p(arr[0]), p(ray_a, C.c_int64), p(arr[1]), p(arr[2]), p(arr[3]), p(arr[4]),
                           p(dn_a), p(vd_a), p(G_a), p(g[0]), p(g[1]), p(g[2]), p(g[3]), p(g[4]), p(g_dn))
    names = ["w", "rgb", "brdf", "brdfj", "vn"]
    for nm, got, leaf in zip(names, g, leaves):
        assert np.allclose(got, leaf.grad.numpy(), rtol=2e-4, atol=2e-5), nm
    if both:
        assert np.allclose(g_dn, leaves[5].grad.numpy(), rtol=2e-4, atol=2e-5)
