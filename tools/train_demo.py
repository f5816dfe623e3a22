"""End-to-end training demonstration on synthetic data: a student field (fresh appearance / MLP / light init) is
trained with the CUDA-graph-replayed step to reproduce a teacher's renderings of the lego scene; prints the loss
curve.  Shows that forward, backward, the static-capacity graph and Adam work together as a training loop."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import __graft_entry__ as g  # noqa: E402
g.build()
from tensoir_b200 import Renderer_TensoIR_train  # noqa: E402
from tensoir_b200.static_step import StaticTrainStep  # noqa: E402
from tensoir_b200.synthetic import make_lego_model, hemisphere_poses, training_batch, n_samples_for  # noqa: E402

grid, batch, steps = 64, 2048, int(sys.argv[1]) if len(sys.argv) > 1 else 150
dev = torch.device("cuda", 0)
teacher = make_lego_model(grid, dev, seed=20211202)
student = make_lego_model(grid, dev, seed=7)             # same geometry, different appearance / MLPs / lights
n_s = n_samples_for(grid)
poses = hemisphere_poses(100)
batches = [training_batch(poses, batch, i) for i in range(8)]
targets = []
with torch.no_grad():
    for rays, li in batches:
        ret = Renderer_TensoIR_train(rays, None, li, teacher, N_samples=n_s, white_bg=True, is_train=False,
                                     is_relight=True, sample_method='fixed_envirmap', device=dev, args=bench.Args)
        targets.append((ret["rgb_map"].clone(), ret["rgb_with_brdf_map"].clone()))
t_rgb, t_brdf = torch.zeros(batch, 3, device=dev), torch.zeros(batch, 3, device=dev)


def loss_fn(ret, m):
    return (torch.mean((ret['rgb_map'] - t_rgb) ** 2) + 0.2 * torch.mean((ret['rgb_with_brdf_map'] - t_brdf) ** 2)
            + 0.0005 * ret['normals_diff_map'].mean() + 0.001 * ret['normals_orientation_loss_map'].mean()
            + 0.001 * ret['roughness_smoothness_loss'] + 0.001 * ret['albedo_smoothness_loss'] + 4e-5 * m.density_L1())


opt = torch.optim.Adam(student.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99), fused=True, capturable=True)
st = StaticTrainStep(student, opt, batch, n_s, bench.Args, loss_fn, device=dev)
st.calibrate(batches)
t_rgb.copy_(targets[0][0]); t_brdf.copy_(targets[0][1])
st.capture(warmup=2)
curve = []
for it in range(steps):
    k = it % len(batches)
    t_rgb.copy_(targets[k][0]); t_brdf.copy_(targets[k][1])
    loss = st.run(*batches[k])
    if it % 10 == 0 or it == steps - 1:
        curve.append((it, float(loss.detach())))
print(json.dumps({"what": "student trained against teacher renderings, cuda-graph step", "grid": grid, "batch": batch,
                  "loss_curve": curve, "overflowed_steps": st.overflowed()}))
assert curve[-1][1] < 0.7 * curve[0][1], "loss did not decrease"
