"""GPU check of the fused primary tail: build tail.cu, run forward + backward on random CUDA data and compare with
torch autograd of the expressions in tensoir_b200/primary.py (the same comparison test_tail_host.py makes on the CPU).

    python experiments/primary_tail/check_gpu.py        # exit code 0 = match
"""
import os
import subprocess
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
so = os.path.join(HERE, "libtir_tail.so")
if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(os.path.join(HERE, f))
                                                       for f in ("tail.cu", "tail_body.h")):
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-shared",
                    "-Xcompiler", "-fPIC", "-o", so, os.path.join(HERE, "tail.cu")], check=True)
from tail_op import fused_tail                      # noqa: E402
from test_tail_host import torch_tail               # noqa: E402

dev = "cuda:0"
bad = 0
for both in (True, False):
    torch.manual_seed(3)
    n, n_rays = 20000, 4096
    ray = torch.sort(torch.randint(0, n_rays, (n,)))[0].to(dev)
    w = (torch.rand(n) * 0.3).to(dev)
    w[::17] = 0.0
    rgb, brdf, brdfj = torch.rand(n, 3, device=dev), torch.rand(n, 4, device=dev), torch.rand(n, 4, device=dev)
    vn = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)
    dn = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1) if both else None
    vd = torch.nn.functional.normalize(torch.randn(n_rays, 3, device=dev), dim=-1)
    G = torch.randn(n_rays, 14, device=dev)

    def run(fn):
        leaves = [t.clone().requires_grad_(True) for t in (w, rgb, brdf, brdfj, vn)] + \
                 ([dn.clone().requires_grad_(True)] if both else [])
        out = fn(leaves[0], ray, leaves[1], leaves[2], leaves[3], leaves[4], leaves[5] if both else None, vd, n_rays)
        out.backward(G)
        return out.detach(), [l.grad for l in leaves]

    def ref(w_, ray_, rgb_, brdf_, brdfj_, vn_, dn_, vd_, n_rays_):
        packed = torch_tail(w_.cpu(), ray_.cpu(), rgb_.cpu(), brdf_.cpu(), brdfj_.cpu(), vn_.cpu(),
                            None if dn_ is None else dn_.cpu(), vd_.cpu(), n_rays_)
        return packed.to(dev)
    got, g_got = run(fused_tail)
    want, g_want = run(ref)
    errs = [float((got - want).abs().max())] + [float((a - b).abs().max()) for a, b in zip(g_got, g_want)]
    ok = errs[0] < 2e-4 and all(e < 2e-4 for e in errs[1:])
    bad += not ok
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        run(fused_tail)
    torch.cuda.synchronize()
    print(f"both_normals={both}: max|err| forward {errs[0]:.2e}, grads {max(errs[1:]):.2e}  {'OK' if ok else 'MISMATCH'}"
          f"   ({(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per fwd+bwd incl. host)")
sys.exit(bad)
