"""python -m pytest experiments/primary_epilogue/test_epilogue_host.py -q   (CPU)
epilogue_body.h (host build) against torch autograd of the per-ray tail of tensoir_b200/primary.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
# column layout of the [n, 18] output (= struct EpiOut): rgb 0:3, depth 3, normal 4:7, albedo 7:10, rough 10,
# fresnel 11:14, normals_diff 14, orientation 15, albedo cost 16, roughness cost 17
N_OUT = 18


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("epi") / "libepi_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                    os.path.join(HERE, "epilogue_host.cpp")], check=True)
    return C.CDLL(so)


def srgb(t):
    t = torch.clamp(t, min=0, max=1)
    return torch.where(t <= 0.0031308, t * 12.92, 1.055 * torch.pow(t + 1e-6, 1 / 2.4) - (1.055 - 1))


def torch_epilogue(packed, acc, depth, rays, fresnel0, bg):
    rgb, normal, albedo = packed[:, 0:3], packed[:, 3:6], packed[:, 6:9]
    rough, ac, rc, nd, no = packed[:, 9:10], packed[:, 10:11], packed[:, 11:12], packed[:, 12:13], packed[:, 13:14]
    fres = torch.zeros_like(albedo).fill_(fresnel0)
    if bg:
        depth = depth + (1. - acc) * rays[..., -1]
        rgb = rgb + (1. - acc[..., None])
        normal = normal + (1 - acc[..., None]) * torch.tensor([0., 0., 1.])
        albedo = albedo + (1 - acc[..., None])
        rough = rough + (1 - acc[..., None])
        fres = fres + (1 - acc[..., None])
    rgb = srgb(rgb.clamp(0, 1))
    return torch.cat([rgb, depth[:, None], F.normalize(normal, p=2, dim=-1, eps=1e-6), albedo.clamp(0, 1),
                      rough.clamp(0, 1), fres.clamp(0, 1), nd, no, ac, rc], dim=-1)


def p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@pytest.mark.parametrize("bg", [1, 0])
def test_epilogue_matches_autograd(lib, bg):
    torch.manual_seed(9)
    n = 300
    packed = torch.rand(n, 14) * 1.3 - 0.1                     # some channels outside [0,1]: clamp gates
    packed[:20, 0:3] *= 0.002                                  # linear branch of the sRGB curve
    packed[20:25, 3:6] = 0.0                                   # degenerate normals (norm < eps when bg = 0)
    acc = torch.rand(n)
    acc[25:40] = 1.0
    depth, rays = torch.rand(n) * 4, torch.randn(n, 6)
    leaves = [t.clone().requires_grad_(True) for t in (packed, acc, depth)]
    want = torch_epilogue(leaves[0], leaves[1], leaves[2], rays, 0.04, bg)
    G = torch.randn(n, N_OUT)
    want.backward(G)
    a = [t.numpy().astype(np.float32) for t in (packed, acc, depth, rays)]
    out = np.zeros((n, N_OUT), np.float32)
    lib.epilogue_forward_host(C.c_int64(n), p(a[0]), p(a[1]), p(a[2]), p(a[3]), C.c_float(0.04), C.c_int(bg), p(out))
    assert np.allclose(out, want.detach().numpy(), rtol=2e-5, atol=2e-6)
    Ga = G.numpy().astype(np.float32)
    gp, ga, gd = np.zeros((n, 14), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    lib.epilogue_backward_host(C.c_int64(n), p(a[0]), p(a[1]), p(a[2]), p(a[3]), C.c_float(0.04), C.c_int(bg), p(Ga),
                               p(gp), p(ga), p(gd))
    ok_rows = torch.ones(n, dtype=torch.bool)
    if not bg:
        ok_rows[20:25] = False        # exact-zero normals: torch propagates 0/0 conventions of norm(); kernel uses g/eps
    m = ok_rows.numpy()
    assert np.allclose(gp[m], leaves[0].grad.numpy()[m], rtol=3e-4, atol=3e-5)
    assert np.allclose(ga[m], (leaves[1].grad.numpy() if bg else np.zeros(n, np.float32))[m], rtol=3e-4, atol=3e-4)
    assert np.allclose(gd, leaves[2].grad.numpy(), rtol=1e-6)
