"""Whole-step CUDA-graph capture of the relight training step.

Eagerly, one step is ~1000 kernel launches (compacted-list bookkeeping, MLP heads, shading epilogue, autograd,
Adam) and is bound by host launch latency, not by the GPU.  Here every dynamically-sized list of the primary march is
given a static capacity (rows past the real total are zero padding; overflow is counted on the device and checked by
the host), host-drawn random numbers are written into static device buffers before each replay, and the whole
forward + loss + backward + (optional gradient all-reduce) + Adam is captured once into a CUDA graph and replayed.

Anything the host changes from step to step must live in device memory to reach a replay: ray batch / light indices
(copied in by ``run``), the host-drawn randoms (staged by ``run``), learning rates (``lr_tensors`` + ``scale_lr``), and
any loss weight that ``loss_fn`` varies over the iterations (the reference decays its TV weights,
train_tensoIR.py:277-282) — keep those in 0-dim device tensors and update them in place between replays; Python floats
are baked in at capture.  Grid maintenance (updateAlphaMask / shrink / upsample) needs a new capture (``run`` checks).

Numerics are those of the eager path except for (i) the device-side xyz-noise draw, which has the padded shape and
therefore a different random stream, and (ii) atomics ordering.  Parity tests run the eager path.
"""
from __future__ import annotations

import torch

from . import _lib
from .device_field import mark_parameters_updated
from .renderer import Renderer_TensoIR_train


def lr_tensors(param_groups, device):
    """Optimizer param groups with their learning rates turned into 0-dim fp32 tensors on ``device`` (what
    ``torch.optim.Adam(..., capturable=True)`` reads on the device at every replay), so that learning-rate decay keeps
    working under CUDA-graph replay (:meth:`StaticTrainStep.scale_lr`)."""
    out = []
    for grp in param_groups:
        grp = dict(grp)
        grp["lr"] = torch.as_tensor(float(grp["lr"]), dtype=torch.float32, device=device)
        out.append(grp)
    return out


class StaticTrainStep:
    def __init__(self, model, optimizer, n_rays, n_samples, args, loss_fn, *, sample_method="stratified_sampling",
                 cap_valid=None, cap_app=None, grad_bucket=None, device=None):
        self.model, self.opt, self.args, self.loss_fn = model, optimizer, args, loss_fn
        self.n_rays, self.n_samples, self.sample_method = int(n_rays), int(n_samples), sample_method
        self.bucket = grad_bucket
        dev = device or model.device
        self.dev = dev
        self.rays = torch.zeros(n_rays, 6, device=dev)
        self.light_idx = torch.zeros(n_rays, 1, dtype=torch.int32, device=dev)
        self.static = {"cap_valid": int(cap_valid or 256 * n_rays), "cap_app": int(cap_app or 16 * n_rays),
                       "jitter": torch.zeros(n_rays, 1, device=dev),
                       "dirs": torch.zeros(model.envmap_h * model.envmap_w, 3, device=dev),
                       "overflow": torch.zeros((), dtype=torch.int64, device=dev)}
        self.graph = None
        self.loss = None
        self.launches_per_replay = 0
        self._pin_j = torch.zeros(n_rays, 1).pin_memory()
        self._pin_d = torch.zeros(model.envmap_h * model.envmap_w, 3).pin_memory()

    @torch.no_grad()
    def calibrate(self, batches, headroom=1.2):
        """Size the static lists from eager marches over a few representative batches (max count x headroom)."""
        from . import primary
        self.model.__dict__.pop("_tir_static", None)
        nv = na = nh = 0
        for rays, _ in batches:
            m = primary.march(self.model, rays.to(self.dev).float(), True, self.n_samples)
            nv = max(nv, int(m["xn"].shape[0]))
            na = max(na, int((m["weight"] > self.model.rayMarch_weight_thres).sum().item()))
            nh = max(nh, int((m["acc"] > 0.5).sum().item()))
        self.static["cap_valid"] = int(headroom * nv) + 4096
        self.static["cap_app"] = int(headroom * na) + 1024
        self.static["cap_hit"] = min(self.n_rays, int(headroom * nh) + 64)
        return self.static["cap_valid"], self.static["cap_app"], self.static["cap_hit"]

    # -- host side of one step: the reference's CPU draws, in the reference's order, into the static buffers
    def _stage_host_randoms(self):
        self._pin_j.copy_(torch.rand(self.n_rays, 1))                                    # tensorBase:714-718
        self._pin_d.copy_(self.model.gen_light_incident_dirs(method=self.sample_method))  # tensorBase:511-526
        self.static["jitter"].copy_(self._pin_j, non_blocking=True)
        self.static["dirs"].copy_(self._pin_d, non_blocking=True)

    def _body(self):
        ret = Renderer_TensoIR_train(self.rays, None, self.light_idx, self.model, N_samples=self.n_samples,
                                     white_bg=True, is_train=True, is_relight=True, sample_method=self.sample_method,
                                     chunk_size=160000, device=self.dev, args=self.args)
        loss = self.loss_fn(ret, self.model)
        loss.backward()
        if self.bucket is not None:
            self.bucket.all_reduce_mean()
        self.opt.step()
        return loss

    def capture(self, warmup=3, capture_error_mode="global"):
        """Warm up eagerly on a side stream (allocations, cudaFuncSetAttribute, cuBLAS handles), then capture."""
        self.model.__dict__["_tir_static"] = self.static
        _lib.load()
        # warm-up and capture run on the SAME side stream: autograd's AccumulateGrad nodes remember the stream they
        # were created on, and a mismatch with the capturing stream would need a cross-stream sync that is illegal
        # during capture
        s = torch.cuda.Stream()
        self._stream = s
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._stage_host_randoms()
                self.opt.zero_grad(set_to_none=True)
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._stage_host_randoms()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        l0 = _lib.launch_count
        with torch.cuda.graph(self.graph, stream=s, capture_error_mode=capture_error_mode):
            self.loss = self._body()
        self.launches_per_replay = _lib.launch_count - l0
        torch.cuda.synchronize()
        # The captured kernels hold raw pointers into buffers owned elsewhere (the sample-list scratch, the alpha-mask
        # and VM shadows).  Keep references so that a later eager call that re-sizes / rebuilds them cannot hand the
        # memory back to the allocator while this graph can still be replayed.
        df = self.model.__dict__.get("_tir_device_field")
        self._keepalive = [self.model.__dict__.get("_tir_scratch"), self.model.alphaMask,
                           None if df is None else (df.dplane, df.dline, df.aplane, df.aline, df.amask, df.acell)]
        self._captured_for = self._model_signature()
        return self

    def _model_signature(self):
        """Identity of everything the graph baked in by address: grid maintenance (updateAlphaMask / shrink /
        upsample_volume_grid, train_tensoIR.py:386-422) replaces these objects and requires a new capture."""
        m = self.model
        vm = list(m.density_plane) + list(m.density_line) + list(m.app_plane) + list(m.app_line)
        return (id(m.alphaMask), tuple((p.data_ptr(), tuple(p.shape)) for p in vm))

    def run(self, rays, light_idx):
        """One training step.  ``rays`` / ``light_idx`` may live on the host (pinned) or on the device."""
        if self.graph is None:
            raise RuntimeError("StaticTrainStep.run() before capture()")
        if self._model_signature() != self._captured_for:
            raise RuntimeError("the model's alpha mask or VM factors were replaced (updateAlphaMask / shrink / "
                               "upsample_volume_grid) after capture(): build and capture a new StaticTrainStep")
        self.rays.copy_(rays, non_blocking=True)
        self.light_idx.copy_(light_idx.reshape(-1, 1), non_blocking=True)
        self._stage_host_randoms()
        self.graph.replay()
        mark_parameters_updated()        # the replayed optimizer step is invisible to the host-side bookkeeping
        _lib.launch_count += self.launches_per_replay
        return self.loss

    def scale_lr(self, factor: float):
        """The train loop's per-iteration decay ``param_group['lr'] = param_group['lr'] * lr_factor``
        (train_tensoIR.py:380-381) for a captured step: a Python-float learning rate is baked into the graph at capture,
        so the rates must be 0-dim device tensors (see :func:`lr_tensors`) and are scaled IN PLACE here, between
        replays, where the next replay reads them."""
        for grp in self.opt.param_groups:
            lr = grp["lr"]
            if not isinstance(lr, torch.Tensor):
                raise TypeError("scale_lr() needs tensor learning rates: build the optimizer from "
                                "static_step.lr_tensors(model.get_optparam_groups(...), device)")
            lr.mul_(factor)

    def overflowed(self) -> int:
        """Number of steps whose lists exceeded the static capacities (host sync; call occasionally)."""
        return int(self.static["overflow"].item())

    def release(self):
        self.model.__dict__.pop("_tir_static", None)
