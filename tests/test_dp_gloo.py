"""Data-parallel plumbing on CPU: world_size-2 gloo, the flattened-bucket all-reduce averages gradients and keeps
replicas identical; sharding is disjoint and complete."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tensoir_b200.dp import GradBucket, shard_batch, broadcast_parameters
    torch.manual_seed(1234 + rank)                       # different init per rank ...
    net = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.ReLU(), torch.nn.Linear(8, 3))
    broadcast_parameters(net.parameters())               # ... made identical by the broadcast
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(16, 6, generator=g), torch.randn(16, 3, generator=g)
    lo, hi = shard_batch(16, rank, world)
    bucket = GradBucket(net.parameters())
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    for _ in range(3):
        opt.zero_grad(set_to_none=False)
        loss = ((net(x[lo:hi]) - y[lo:hi]) ** 2).mean()
        loss.backward()
        bucket.all_reduce_mean()
        opt.step()
    out[rank] = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    if rank == 0:
        # single-process reference on the full batch
        torch.manual_seed(1234)
        ref = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.ReLU(), torch.nn.Linear(8, 3))
        opt = torch.optim.Adam(ref.parameters(), lr=1e-2)
        for _ in range(3):
            opt.zero_grad()
            ((ref(x) - y) ** 2).mean().backward()
            opt.step()
        out["ref"] = torch.cat([p.detach().reshape(-1) for p in ref.parameters()])
    dist.destroy_process_group()


def test_bucket_allreduce_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.equal(out[0], out[1])                                   # replicas stay identical
    assert torch.allclose(out[0], out["ref"], rtol=1e-5, atol=1e-6)      # mean of shard grads == full-batch grad


def test_shard_batch_partition():
    from tensoir_b200.dp import shard_batch
    cover = []
    for r in range(8):
        lo, hi = shard_batch(4096, r, 8)
        cover += list(range(lo, hi))
    assert cover == list(range(4096))


def _view_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tensoir_b200 import relight
    from tensoir_b200.dp import shard_views
    mine = shard_views(7)
    # the sharded relighting loop with the per-view renderer stubbed out (host logic only: sharding + metric gather)
    relight.relight_view = lambda model, env, names, rays, **kw: {n: (rays[:, :3] * (1 + k), None)
                                                                  for k, n in enumerate(names)}
    views = [torch.full((4, 6), float(v)) for v in range(7)]
    local, metrics = relight.relight_views_sharded(None, None, ["a", "b"], views)
    out[rank] = (mine, sorted(local), metrics)
    dist.destroy_process_group()


def test_view_sharding_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_view_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    (m0, l0, met0), (m1, l1, met1) = out[0], out[1]
    assert m0 == [0, 2, 4, 6] and m1 == [1, 3, 5] and l0 == m0 and l1 == m1     # disjoint, complete, round-robin
    assert met0 == met1 and sorted(met0) == list(range(7))                         # every rank holds every metric
    assert met0[3] == {"a": 3.0, "b": 6.0}
