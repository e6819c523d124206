"""Per-stage device times of one config through the public API (f3dgs_profile_*).  usage: stage_times.py cfg [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "feature-3dgs_b200"))
import torch, scenegen
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer, _C
name = sys.argv[1]; iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
sc = scenegen.make_config(name); cam = sc.cameras[0]; dev = "cuda"
t = scenegen.to_torch(sc, dev, requires_grad=True)
rs = GaussianRasterizationSettings(**scenegen.settings_kwargs(sc, cam, dev))
gc, gf, gd = [torch.from_numpy(g).to(dev) for g in scenegen.upstream_grads(cam.image_height, cam.image_width, sc.C)]
names = ["pre_fwd", "scan", "dup", "sort", "ranges", "comp_fwd", "comp_bwd", "pre_bwd"]
for it in range(iters + 2):
    if it == 2:
        torch.cuda.synchronize(); _C.profile_read(); _C.profile_enable(True)
    m2 = torch.zeros_like(t["means3D"], requires_grad=True)
    color, feat, radii, depth = GaussianRasterizer(rs)(means3D=t["means3D"], means2D=m2, opacities=t["opacities"], shs=t["shs"],
        semantic_feature=t["semantic_feature"] if sc.C else None, scales=t["scales"], rotations=t["rotations"])
    outs, gos = [color, depth], [gc, gd]
    if sc.C: outs.append(feat); gos.append(gf)
    torch.autograd.backward(outs, gos)
    for k in t: t[k].grad = None
torch.cuda.synchronize()
ms, cnt = _C.profile_read()
print(name, "BPA", os.environ.get("F3DGS_BPA", "default"), {n: round(m / max(c, 1), 3) for n, m, c in zip(names, ms, cnt)})
