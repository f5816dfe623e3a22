"""torch.profiler breakdown of one training step (which kernels / how much host overhead)."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=300)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
import __graft_entry__ as g  # noqa: E402
g.build()
from tensoir_b200 import Renderer_TensoIR_train  # noqa: E402
from tensoir_b200.synthetic import make_lego_model, hemisphere_poses, training_batch, n_samples_for  # noqa: E402

dev = torch.device("cuda:0")
model = make_lego_model(a.grid, dev)
opt = torch.optim.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99), fused=True)
poses = hemisphere_poses(100)
n_s = n_samples_for(a.grid)
target = torch.full((4096, 3), 0.5, device=dev)
batches = [tuple(t.to(dev) for t in training_batch(poses, 4096, i)) for i in range(a.steps + 3)]


def step(rays, li):
    ret = Renderer_TensoIR_train(rays, None, li, model, N_samples=n_s, white_bg=True, is_train=True, is_relight=True,
                                 sample_method='stratified_sampling', device=dev, args=bench.Args)
    loss = bench.loss_of(ret, target, model)
    opt.zero_grad(set_to_none=False)
    loss.backward()
    opt.step()


for b in batches[:3]:
    step(*b)
torch.cuda.synchronize()
t0 = time.perf_counter()
for b in batches[3:]:
    step(*b)
torch.cuda.synchronize()
print("wall ms/step", 1e3 * (time.perf_counter() - t0) / a.steps)
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for b in batches[3:]:
        step(*b)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=60))
