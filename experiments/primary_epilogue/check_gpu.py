"""GPU check of the per-ray epilogue: build epilogue.cu, compare forward + backward with torch autograd on CUDA data
(the comparison test_epilogue_host.py makes on the CPU).

    python experiments/primary_epilogue/check_gpu.py        # exit code 0 = match
"""
import ctypes as C
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
so = os.path.join(HERE, "libtir_epilogue.so")
if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(os.path.join(HERE, f))
                                                       for f in ("epilogue.cu", "epilogue_body.h")):
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-shared",
                    "-Xcompiler", "-fPIC", "-o", so, os.path.join(HERE, "epilogue.cu")], check=True)
lib = C.CDLL(so)
from test_epilogue_host import torch_epilogue, N_OUT          # noqa: E402

dev = "cuda:0"


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


bad = 0
for bg in (1, 0):
    torch.manual_seed(9)
    n = 4096
    packed = (torch.rand(n, 14) * 1.3 - 0.1)
    packed[:200, 0:3] *= 0.002
    acc = torch.rand(n)
    acc[250:400] = 1.0
    depth, rays = torch.rand(n) * 4, torch.randn(n, 6)
    leaves = [t.clone().requires_grad_(True) for t in (packed, acc, depth)]
    want = torch_epilogue(leaves[0], leaves[1], leaves[2], rays, 0.04, bg)
    G = torch.randn(n, N_OUT)
    G[:, 16:] = 0.0                                   # the two scalar losses are returned separately by the kernel
    g_loss = torch.tensor([0.7, -0.3])
    (want * G).sum().add(want[:, 16].mean() * g_loss[0] + want[:, 17].mean() * g_loss[1]).backward()

    d = [t.to(dev).contiguous() for t in (packed, acc, depth, rays, G, g_loss)]
    out = torch.zeros(n, N_OUT, device=dev)
    mask = torch.zeros(n, dtype=torch.uint8, device=dev)
    losses = torch.zeros(2, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.tir_epilogue_fwd(C.c_int64(n), ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[3]), C.c_float(0.04), C.c_int(bg),
                              ptr(out), ptr(mask), ptr(losses), st)
    gp, ga, gd = torch.zeros(n, 14, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    rc |= lib.tir_epilogue_bwd(C.c_int64(n), ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[3]), C.c_float(0.04), C.c_int(bg),
                               ptr(d[4]), ptr(d[5]), ptr(gp), ptr(ga), ptr(gd), st)
    torch.cuda.synchronize()
    e_fwd = float((out.cpu() - want.detach()).abs().max())
    e_loss = float((losses.cpu() - torch.stack([want[:, 16].mean(), want[:, 17].mean()]).detach()).abs().max())
    e_mask = int((mask.cpu().bool() != (acc > 0.5)).sum())
    e_gp = float((gp.cpu() - leaves[0].grad).abs().max())
    e_ga = float((ga.cpu() - (leaves[1].grad if leaves[1].grad is not None else torch.zeros(n))).abs().max())
    e_gd = float((gd.cpu() - leaves[2].grad).abs().max())
    ok = rc == 0 and e_fwd < 1e-4 and e_loss < 1e-5 and e_mask == 0 and e_gp < 2e-3 and e_ga < 2e-3 and e_gd < 1e-5
    bad += not ok
    print(f"bg={bg}: rc {rc}, forward {e_fwd:.2e}, losses {e_loss:.2e}, mask mismatches {e_mask}, g_packed {e_gp:.2e}, "
          f"g_acc {e_ga:.2e}, g_depth {e_gd:.2e}  {'OK' if ok else 'MISMATCH'}")
sys.exit(bad)
