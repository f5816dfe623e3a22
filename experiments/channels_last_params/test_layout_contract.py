"""python -m pytest experiments/channels_last_params/test_layout_contract.py -q   (CPU)

Round-2 plan: keep the VM factors in channel-last STORAGE and let the nn.Parameter be its NCHW-shaped view
(`torch.channels_last` strides).  Then the kernels read the parameter storage itself — no shadow copies, no per-step
repack, nothing that can go stale — and the channel-last gradient buffers the scatter kernels fill are accepted by
autograd without a copy.  This file pins the PyTorch behaviours that plan relies on (checked here on the CPU; the CUDA
fused Adam applies the same same-layout rule to its tensor lists)."""
import io

import torch

C, H, W = 16, 9, 7


def _param():
    storage = torch.randn(H, W, C)                                   # what the kernels index: [H][W][C]
    return storage, torch.nn.Parameter(storage.permute(2, 0, 1).unsqueeze(0))


def test_parameter_is_a_view_of_channel_last_storage():
    storage, p = _param()
    assert p.shape == (1, C, H, W) and p.data_ptr() == storage.data_ptr()
    assert p.is_contiguous(memory_format=torch.channels_last)
    line = torch.nn.Parameter(torch.randn(1, C, 11, 1).contiguous(memory_format=torch.channels_last))
    assert line.stride()[1] == 1 and line.stride()[2] == C          # [G][C]


def test_channel_last_gradients_are_taken_without_a_copy():
    _, p = _param()

    class Scatter(torch.autograd.Function):                          # stands for a tir_vm_*_bwd kernel
        @staticmethod
        def forward(ctx, x):
            return x.sum()

        @staticmethod
        def backward(ctx, g):
            buf = torch.ones(H, W, C)
            Scatter.ptr = buf.data_ptr()
            return buf.permute(2, 0, 1).unsqueeze(0)
    Scatter.apply(p).backward()
    assert p.grad.data_ptr() == Scatter.ptr and p.grad.stride() == p.stride()     # layout contract: stolen, not copied


def test_adam_state_and_checkpoints_keep_the_layout():
    _, p = _param()
    p.grad = torch.ones_like(p)                                      # preserve_format: channel-last too
    for fused in (False, True):
        opt = torch.optim.Adam([p], lr=0.02, betas=(0.9, 0.99), fused=fused)
        opt.step()
        assert opt.state[p]["exp_avg"].stride() == p.stride()
    blob = io.BytesIO()
    torch.save({"w": p.detach()}, blob)
    blob.seek(0)
    q = torch.load(blob)["w"]
    assert q.stride() == p.stride() and torch.equal(q, p.detach())
    m = torch.nn.Module()
    m.w = torch.nn.Parameter(torch.zeros(1, C, H, W).contiguous(memory_format=torch.channels_last))
    m.load_state_dict({"w": torch.randn(1, C, H, W)})                # a reference (NCHW) checkpoint loads into it
    assert m.w.is_contiguous(memory_format=torch.channels_last)


def test_upsample_needs_an_explicit_relayout():
    _, p = _param()
    up = torch.nn.functional.interpolate(p.data, size=(12, 10), mode="bilinear", align_corners=True)
    assert not up.is_contiguous(memory_format=torch.channels_last)  # upsample_volume_grid must convert the result
    assert up.contiguous(memory_format=torch.channels_last).stride()[1] == 1
