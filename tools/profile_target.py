"""Short, deterministic launch sequence for ncu: the two secondary kernels on one 4096-ray batch."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=300)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
import __graft_entry__ as g   # noqa: E402
g.build()
from tensoir_b200 import ops  # noqa: E402
from tensoir_b200.synthetic import make_lego_model, hemisphere_poses, training_batch, n_samples_for  # noqa: E402

dev = torch.device("cuda:0")
m = make_lego_model(a.grid, dev)
rays, li = training_batch(hemisphere_poses(100), 4096, 0)
rays, li = rays.to(dev), li.to(dev)
with torch.no_grad():
    out = m(rays, li, is_train=False, is_relight=True, N_samples=n_samples_for(a.grid))
mask = out[9]
surf = (rays[:, :3] + out[1][:, None] * rays[:, 3:])[mask]
torch.manual_seed(0)
dirs = m.gen_light_incident_dirs(method='stratified_sampling').to(dev)
st = ops.SecondaryStages(m, surf, out[2][mask], li[mask], dirs)
for _ in range(a.reps):
    st.march()
    st.mlp()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    st.march()
e1.record()
torch.cuda.synchronize()
print("march_ms", e0.elapsed_time(e1) / 20, "unroll", os.environ.get("TIR_MARCH_UNROLL"), "coarse", os.environ.get("TIR_MARCH_COARSE"))
print("counters", ops.counters_dict(st.counters), "pts", surf.shape[0])
