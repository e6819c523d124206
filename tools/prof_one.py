"""Minimal driver for ncu: a few forward+backward passes of one config through the public API.
usage: python tools/prof_one.py <config> [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "feature-3dgs_b200"))
import torch  # noqa: E402

import scenegen  # noqa: E402
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = scenegen.make_config(name)
cam = sc.cameras[0]
dev = "cuda"
t = scenegen.to_torch(sc, dev, requires_grad=True)
rs = GaussianRasterizationSettings(**scenegen.settings_kwargs(sc, cam, dev))
gc, gf, gd = [torch.from_numpy(g).to(dev) for g in scenegen.upstream_grads(cam.image_height, cam.image_width, sc.C)]
for it in range(iters):
    m2 = torch.zeros_like(t["means3D"], requires_grad=True)
    color, feat, radii, depth = GaussianRasterizer(rs)(
        means3D=t["means3D"], means2D=m2, opacities=t["opacities"], shs=t["shs"],
        semantic_feature=t["semantic_feature"] if sc.C else None, scales=t["scales"], rotations=t["rotations"])
    outs, gos = [color, depth], [gc, gd]
    if sc.C:
        outs.append(feat)
        gos.append(gf)
    torch.autograd.backward(outs, gos)
    for k in t:
        t[k].grad = None
torch.cuda.synchronize()
print("done", name, iters)
