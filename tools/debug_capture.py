import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import __graft_entry__ as g
g.build()
from tensoir_b200.static_step import StaticTrainStep
from tensoir_b200.synthetic import make_lego_model, hemisphere_poses, training_batch, n_samples_for
grid = int(sys.argv[1]); batch = int(sys.argv[2]); variant = sys.argv[3]
if os.environ.get("SETDEV"):
    torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
if os.environ.get("LOADLIB"):
    from tensoir_b200 import Renderer_TensoIR_train, _lib, ops
    from tensoir_b200.dp import GradBucket, broadcast_parameters
    _lib.load()
model = make_lego_model(grid, dev)
if os.environ.get("PARAMS"):
    from tensoir_b200.dp import broadcast_parameters
    broadcast_parameters(model.parameters())
    params = [p for grp in model.get_optparam_groups(0.02, 0.001) for p in (grp["params"] if isinstance(
        grp["params"], (list, tuple, torch.nn.ParameterList)) else list(grp["params"]))]
opt = torch.optim.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99), fused=True, capturable=True)
poses = hemisphere_poses(100)
hb = [training_batch(poses, batch, i) for i in range(26 if os.environ.get("MANYB") else 3)]
if os.environ.get("MANYB"):
    pinned_all = [(r.pin_memory(), l.pin_memory()) for r, l in hb]
target = torch.full((batch, 3), 0.5, device=dev)
def loss_fn(ret, m):
    if variant == "rgb":
        return torch.mean((ret['rgb_map'] - target) ** 2)
    if variant == "rgb_brdf":
        return torch.mean((ret['rgb_map'] - target) ** 2) + 0.2 * torch.mean((ret['rgb_with_brdf_map'] - target) ** 2)
    if variant == "brdf_nograd_normal":
        return torch.mean((ret['rgb_map'] - target) ** 2) + 0.2 * torch.mean((ret['rgb_with_brdf_map'].detach() - target) ** 2)
    if variant == "no_l1":
        return bench.loss_of(ret, target, m) - 4e-5 * m.density_L1()
    return bench.loss_of(ret, target, m)
if os.environ.get("COUNTERS"):
    from tensoir_b200 import ops
    model.__dict__["_tir_counters"] = ops.new_counters(dev)
if os.environ.get("PIN"):
    pinned = [(r.pin_memory(), l.pin_memory()) for r, l in hb]
st = StaticTrainStep(model, opt, batch, n_samples_for(grid), bench.Args, loss_fn, device=dev)
print("caps", st.calibrate(hb))
try:
    st.capture(warmup=int(os.environ.get("WARM", "2")), capture_error_mode=os.environ.get("CEM", "global"))
    l = st.run(hb[0][0], hb[0][1])
    torch.cuda.synchronize()
    print("OK", grid, batch, variant, float(l), "overflow", st.overflowed())
except Exception as e:
    print("FAIL", grid, batch, variant, str(e).split("\n")[0])
