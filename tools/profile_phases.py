"""Per-phase GPU time of the (eager, shape-static) training step: forward phases by record_function range, backward
by autograd node."""
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
model.__dict__["_tir_static"] = st.static
def eager_step(r, l):
    st.rays.copy_(r); st.light_idx.copy_(l.reshape(-1, 1)); st._stage_host_randoms()
    opt.zero_grad(set_to_none=True)
    st._body()
for r, l in hb[:3]:
    eager_step(r, l)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for r, l in hb[3:8]:
        eager_step(r, l)
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = [(e.key, e.device_time_total / 5e3, e.self_device_time_total / 5e3, e.count / 5) for e in ka]
print("---- ranges / autograd nodes (CUDA total ms per step)")
for k, tot, selft, c in sorted(rows, key=lambda x: -x[1]):
    if k.startswith("tir::") or k.startswith("autograd::engine::evaluate_function") or k.startswith("Optimizer"):
        print(f"{tot:8.3f} ms  n={c:6.1f}  {k[:90]}")
print("---- total self CUDA ms/step", sum(r[2] for r in rows))
