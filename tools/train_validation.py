"""Training validation (not just steps): a student field is trained for N iterations against a teacher's renderings of
the lego scene with the schedule the train scripts use — per-iteration learning-rate decay, one updateAlphaMask and
one upsample_volume_grid (both followed by a new optimizer / a re-captured step, train_tensoIR.py:380-422) — once with
the CUDA-graph-replayed step (StaticTrainStep + FusedAdam) and once eagerly (torch.optim.Adam), same batches, targets
and host randoms.  PSNR against the teacher on a held-out batch must rise and the two runs must track each other.

    python tools/train_validation.py [steps=500] > profiles/r2_train_validation.json
"""
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
from tensoir_b200.optim import FusedAdam  # noqa: E402
from tensoir_b200.static_step import StaticTrainStep, lr_tensors  # noqa: E402
from tensoir_b200.synthetic import make_lego_model, hemisphere_poses, training_batch, n_samples_for  # noqa: E402

_real_stdout = sys.stdout
sys.stdout = sys.stderr            # model-building prints must not mix with the one JSON line on stdout
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
grid0, grid1, batch = 64, 96, 2048
mask_at, up_at = int(0.3 * steps), int(0.5 * steps)
lr_factor = 0.1 ** (1.0 / steps)
dev = torch.device("cuda", 0)
poses = hemisphere_poses(100)
teacher = make_lego_model(grid1, dev, seed=20211202)
n_batches = 16
batches = [tuple(t.to(dev) for t in training_batch(poses, batch, i)) for i in range(n_batches + 1)]
targets = []
with torch.no_grad():
    for rays, li in batches:
        ret = Renderer_TensoIR_train(rays, None, li, teacher, N_samples=n_samples_for(grid1), white_bg=True,
                                     is_train=False, is_relight=True, sample_method='fixed_envirmap', device=dev,
                                     args=bench.Args)
        targets.append((ret["rgb_map"].clone(), ret["rgb_with_brdf_map"].clone()))
eval_batch, eval_target = batches[-1], targets[-1]
t_rgb, t_brdf = torch.zeros(batch, 3, device=dev), torch.zeros(batch, 3, device=dev)


def loss_fn(ret, m, l1_in_opt):
    loss = (torch.mean((ret['rgb_map'] - t_rgb) ** 2) + 0.2 * torch.mean((ret['rgb_with_brdf_map'] - t_brdf) ** 2)
            + 0.0005 * ret['normals_diff_map'].mean() + 0.001 * ret['normals_orientation_loss_map'].mean()
            + 0.001 * ret['roughness_smoothness_loss'] + 0.001 * ret['albedo_smoothness_loss'])
    return loss if l1_in_opt else loss + 4e-5 * m.density_L1()


def psnr(model, grid):
    with torch.no_grad():
        ret = Renderer_TensoIR_train(eval_batch[0], None, eval_batch[1], model, N_samples=n_samples_for(grid),
                                     white_bg=True, is_train=False, is_relight=True, sample_method='fixed_envirmap',
                                     device=dev, args=bench.Args)
    out = []
    for k, t in (("rgb_map", eval_target[0]), ("rgb_with_brdf_map", eval_target[1])):
        mse = torch.mean((ret[k] - t) ** 2).clamp_min(1e-12)
        out.append(float(-10 * torch.log10(mse)))
    return out


def run(graphed):
    torch.manual_seed(11)
    model = make_lego_model(grid0, dev, seed=7)          # same geometry, different appearance / MLPs / lights
    grid, scale = grid0, 1.0
    curve, events = [], []

    def new_step():
        groups = model.get_optparam_groups(0.02 * scale, 0.001 * scale)
        if graphed:
            opt = FusedAdam(lr_tensors(groups, dev), betas=(0.9, 0.99))
            opt.density_l1(model, 4e-5)
            st = StaticTrainStep(model, opt, batch, n_samples_for(grid), bench.Args,
                                 lambda ret, m: loss_fn(ret, m, True), device=dev)
            st.calibrate(batches[:4])
            st.capture(warmup=1)
            return opt, st
        return torch.optim.Adam(groups, betas=(0.9, 0.99)), None

    opt, st = new_step()
    for it in range(steps):
        k = it % n_batches
        t_rgb.copy_(targets[k][0]); t_brdf.copy_(targets[k][1])
        if graphed:
            loss = st.run(*batches[k])
            st.scale_lr(lr_factor)
        else:
            ret = Renderer_TensoIR_train(batches[k][0], None, batches[k][1], model, N_samples=n_samples_for(grid),
                                         white_bg=True, is_train=True, is_relight=True,
                                         sample_method='stratified_sampling', device=dev, args=bench.Args)
            loss = loss_fn(ret, model, False)
            opt.zero_grad()
            loss.backward()
            opt.step()
            for grp in opt.param_groups:
                grp['lr'] = grp['lr'] * lr_factor
        scale *= lr_factor
        if it % 25 == 0 or it == steps - 1:
            if st is not None:
                st.flush()
            curve.append([it, float(loss.detach())] + psnr(model, grid))
        if it == mask_at or it == up_at:
            if st is not None:
                st.flush()
                events.append({"it": it, "redone": st.redone, "recaptures": st.recaptures, "lists": st.capacities()})
                st.release()
            if it == mask_at:
                model.updateAlphaMask((grid, grid, grid))
            else:
                grid = grid1
                model.upsample_volume_grid([grid] * 3)
            opt, st = new_step()
    if st is not None:
        st.flush()
        events.append({"it": steps, "redone": st.redone, "recaptures": st.recaptures, "lists": st.capacities(),
                       "overflowed_replays": st.overflowed()})
        st.release()
    return curve, events


res = {"what": f"student {grid0}^3 -> {grid1}^3 trained {steps} iterations against teacher renderings; lr x {lr_factor:.5f} "
               f"per iteration, updateAlphaMask at {mask_at}, upsample_volume_grid at {up_at}; columns: iteration, loss, "
               f"PSNR rgb_map, PSNR rgb_with_brdf_map (held-out batch vs teacher)", "batch": batch}
res["graph_replay_fused_adam"], res["graph_events"] = run(True)
res["eager_torch_adam"], _ = run(False)
a, b = res["graph_replay_fused_adam"], res["eager_torch_adam"]
res["psnr_gain_db"] = [a[-1][2] - a[0][2], a[-1][3] - a[0][3]]
res["final_psnr_gap_db"] = [abs(a[-1][2] - b[-1][2]), abs(a[-1][3] - b[-1][3])]
# the two runs consume the host / device random streams differently (jitter, stratified light directions, xyz noise), so
# they are two samples of the same stochastic training, not one trajectory: both must gain, and end within 3 dB
res["ok"] = bool(res["psnr_gain_db"][0] > 3 and res["psnr_gain_db"][1] > 3 and max(res["final_psnr_gap_db"]) < 3.0)
sys.stdout = _real_stdout
print(json.dumps(res))
sys.exit(0 if res["ok"] else 1)
