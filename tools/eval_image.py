"""Full-image evaluation timing (BASELINE config 2, second half): 800x800 view, 4096-ray chunks, is_train=False,
fixed_envirmap, through the unchanged boundary call of renderer.py:225-239."""
import json, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import __graft_entry__ as g
g.build()
from tensoir_b200 import Renderer_TensoIR_train, ops
from tensoir_b200.synthetic import make_lego_model, hemisphere_poses, image_rays
dev = torch.device("cuda", 0)
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 300
model = make_lego_model(grid, dev)
rays = image_rays(hemisphere_poses(200)[7]).to(dev)
li = torch.zeros(rays.shape[0], 1, dtype=torch.int32, device=dev)
cnt = ops.new_counters(dev)
model.__dict__["_tir_counters"] = cnt

def render():
    outs = []
    with torch.no_grad():
        for s in range(0, rays.shape[0], 4096):
            ret = Renderer_TensoIR_train(rays[s:s + 4096], None, li[s:s + 4096], model, N_samples=-1, ndc_ray=False,
                                         white_bg=True, is_train=False, is_relight=True,
                                         sample_method='fixed_envirmap', chunk_size=160000, device=dev, args=bench.Args)
            outs.append(ret["rgb_with_brdf_map"])
    return torch.cat(outs)

render()
torch.cuda.synchronize()
cnt.zero_()
t0 = time.perf_counter()
img = render()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
c = ops.counters_dict(cnt)
print(json.dumps({"what": "800x800 eval view, 4096-ray chunks, fixed_envirmap, eager", "grid": grid, "seconds": dt,
                  "primary_rays_per_s": rays.shape[0] / dt, "all_rays_per_s": c["rays"] / dt, "counters": c,
                  "mean_rgb": float(img.mean())}))
