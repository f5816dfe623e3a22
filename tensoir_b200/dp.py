"""Ray-batch data parallelism: one process per GPU, every rank runs the same model on its own
slice of rays, and the parameter gradients are averaged with ONE all-reduce of a single flattened fp32 bucket
per step (NCCL over NVLink 5 / NVSwitch; gloo in the CPU tests).  The reference itself has no data exchange
(vestigial init_process_group + one barrier, train_tensoIR.py:21-27, utils.py:231-242).

Everything else (regularisers, updateAlphaMask / shrink / upsample, Adam) is a deterministic function of the
identical parameters and is simply replicated, so this is the only collective of the training path.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def _params_with_grad(params: Iterable[torch.nn.Parameter]) -> List[torch.nn.Parameter]:
    return [p for p in params if p.requires_grad]


class GradBucket:
    """Persistent flat fp32 gradient bucket; ``p.grad`` of every parameter becomes a view into it, so the
    all-reduce needs no per-step flatten / unflatten copies.  Rebuild after shrink / upsample (new Parameters)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], extra: int = 1):
        self.params = _params_with_grad(params)
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        # `extra` spare floats at the end ride along with the gradients in the same all-reduce (the CUDA-graph step
        # puts its list-overflow flag there, so that one rank's overflow makes every rank skip the same update)
        self.flat = torch.zeros(n + extra, dtype=torch.float32, device=dev)
        self.extra = self.flat[n:]
        o = 0
        self.views = []
        for p in self.params:
            # same strides as the parameter (channel-last VM factors): fused optimizers require params and grads to
            # share one layout, and autograd's layout contract then never re-lays the gradient out
            chunk = self.flat[o:o + p.numel()]
            dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
            v = chunk.as_strided(p.shape, p.stride()) if dense else chunk.view_as(p)
            self.views.append(v)
            o += p.numel()

    def numel(self):
        return self.flat.numel()

    def gather(self):
        """Copy (or alias) the parameters' grads into the bucket; missing grads count as zero."""
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
                p.grad = v          # every rank must apply the averaged gradient, also one whose batch produced none
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v

    def all_reduce_mean(self, group=None):
        self.gather()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))


def shard_batch(n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of an n-ray batch owned by ``rank`` (equal sizes; n must divide evenly so that
    per-rank loss means average to the global mean)."""
    if n % world:
        raise ValueError(f"batch {n} not divisible by world size {world}")
    per = n // world
    return rank * per, (rank + 1) * per


def broadcast_parameters(params: Iterable[torch.Tensor], src: int = 0, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for p in params:
            t = p.data
            if not t.is_contiguous() and t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
                t = t.permute(0, 2, 3, 1)               # the same storage seen as a contiguous tensor (NCCL needs one)
            dist.broadcast(t, src=src, group=group)


# ---- evaluation / relighting: views are independent, so they are sharded; the only exchange is the metric gather ----
def shard_views(n_views: int, rank: int = None, world: int = None) -> List[int]:
    """View ids rendered by ``rank`` (round-robin, so that ranks finish together when cost varies smoothly along a
    camera path).  Defaults to the initialised process group, or a single process."""
    if rank is None or world is None:
        on = dist.is_available() and dist.is_initialized()
        rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    return list(range(rank, n_views, world))


def gather_view_results(local: dict, group=None) -> dict:
    """Merge the per-view results {view_id: picklable metrics} of all ranks; every rank returns the full dict.
    Small host objects only (PSNR / timing per view) — images stay on the rank that rendered them."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, dict(local), group=group)
    merged = {}
    for part in parts:
        dup = merged.keys() & part.keys()
        if dup:
            raise RuntimeError(f"views rendered by more than one rank: {sorted(dup)}")
        merged.update(part)
    return merged
