"""Kernel-time breakdown of the CUDA-graph replayed training step."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import __graft_entry__ as g
g.build()
from tensoir_b200.static_step import StaticTrainStep
from tensoir_b200.synthetic import make_lego_model, hemisphere_poses, training_batch, n_samples_for
from torch.profiler import profile, ProfilerActivity
grid, batch = 300, 4096
dev = torch.device("cuda", 0)
model = make_lego_model(grid, dev)
opt = torch.optim.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99), fused=True, capturable=True)
poses = hemisphere_poses(100)
hb = [training_batch(poses, batch, i) for i in range(8)]
target = torch.full((batch, 3), 0.5, device=dev)
st = StaticTrainStep(model, opt, batch, n_samples_for(grid), bench.Args, lambda r, m: bench.loss_of(r, target, m), device=dev)
print("caps", st.calibrate(hb[:3]))
st.capture(warmup=3)
for r, l in hb[:3]:
    st.run(r, l)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for r, l in hb[3:8]:
        st.run(r, l)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
