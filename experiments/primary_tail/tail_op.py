"""Autograd wrapper for the fused primary tail (experiments; promote into tensoir_b200/ after a GPU run).

Replaces, in tensoir_b200/primary.py:_forward_relight_tail, the block from ``v_alb, v_rough = ...`` to the 14-channel
``_segment_sum`` (about 30 forward + 60 backward torch kernels) with two launches:

    packed = fused_tail(w_a, r_a, rgb, brdf, brdf_j, v_n, d_n_or_None, viewdirs, n_rays)     # [n_rays, 14]

Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o libtir_tail.so tail.cu
"""
import ctypes as C
import os

import torch

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtir_tail.so"))
    return _LIB


def _p(t):
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("fused_tail needs contiguous CUDA tensors")
    return C.c_void_p(t.data_ptr())


class _Tail(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, ray, rgb, brdf, brdfj, vn, dn, viewdirs, n_rays):
        args = [t.contiguous().float() if t is not None else None for t in (w, rgb, brdf, brdfj, vn, dn)]
        ray, viewdirs = ray.contiguous(), viewdirs.contiguous().float()
        packed = torch.zeros(n_rays, 14, device=w.device)
        rc = _lib().tir_tail_fwd(C.c_int64(w.shape[0]), _p(args[0]), _p(ray), _p(args[1]), _p(args[2]), _p(args[3]),
                                 _p(args[4]), _p(args[5]), _p(viewdirs), _p(packed),
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc:
            raise RuntimeError(f"tir_tail_fwd: {rc}")
        ctx.save_for_backward(ray, viewdirs, *[a for a in args if a is not None])
        ctx.has_dn = dn is not None
        return packed

    @staticmethod
    def backward(ctx, g_packed):
        ray, viewdirs, w, rgb, brdf, brdfj, vn, *rest = ctx.saved_tensors
        dn = rest[0] if ctx.has_dn else None
        g = [torch.empty_like(t) for t in (w, rgb, brdf, brdfj, vn)]
        g_dn = torch.empty_like(dn) if dn is not None else None
        rc = _lib().tir_tail_bwd(C.c_int64(w.shape[0]), _p(w), _p(ray), _p(rgb), _p(brdf), _p(brdfj), _p(vn), _p(dn),
                                 _p(viewdirs), _p(g_packed.contiguous().float()), _p(g[0]), _p(g[1]), _p(g[2]),
                                 _p(g[3]), _p(g[4]), _p(g_dn), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc:
            raise RuntimeError(f"tir_tail_bwd: {rc}")
        return g[0], None, g[1], g[2], g[3], g[4], g_dn, None, None


def fused_tail(w, ray, rgb, brdf, brdfj, vn, dn, viewdirs, n_rays):
    """-> [n_rays, 14] = sum over each ray's samples of w * [rgb, vn, albedo, roughness, albedo cost, roughness cost,
    |vn - dn|^2, clamp(viewdir . vn, 0)] (the last two are 0 when ``dn`` is None)."""
    return _Tail.apply(w, ray.long(), rgb, brdf, brdfj, vn, dn, viewdirs, int(n_rays))
