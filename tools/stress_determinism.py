"""Race hunt: the forward has no atomics, so every output must be bit-identical run to run; the backward's float
atomics may reorder, so its gradients must agree within a small fraction of the parity tolerance.
usage: stress_determinism.py cfg [runs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "feature-3dgs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, scenegen, parity
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
name = sys.argv[1]; runs = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sc = scenegen.make_config(name); cam = sc.cameras[0]; dev = "cuda"
t = scenegen.to_torch(sc, dev, requires_grad=True)
rs = GaussianRasterizationSettings(**scenegen.settings_kwargs(sc, cam, dev))
gc, gf, gd = [torch.from_numpy(g).to(dev) for g in scenegen.upstream_grads(cam.image_height, cam.image_width, sc.C)]
base, bad_fwd, worst = None, 0, 0.0
for it in range(runs):
    m2 = torch.zeros_like(t["means3D"], requires_grad=True)
    color, feat, radii, depth = GaussianRasterizer(rs)(means3D=t["means3D"], means2D=m2, opacities=t["opacities"], shs=t["shs"],
        semantic_feature=t["semantic_feature"] if sc.C else None, scales=t["scales"], rotations=t["rotations"])
    outs, gos = [color, depth], [gc, gd]
    if sc.C: outs.append(feat); gos.append(gf)
    torch.autograd.backward(outs, gos)
    cur = dict(color=color.detach().clone(), feat=feat.detach().clone(), depth=depth.detach().clone(), radii=radii.clone())
    g = {k: t[k].grad.clone() for k in t if t[k].grad is not None}; g["means2D"] = m2.grad.clone()
    for k in t: t[k].grad = None
    if base is None:
        base, gbase = cur, g
        continue
    for k in cur:
        if not torch.equal(cur[k], base[k]):
            bad_fwd += 1
            d = (cur[k] != base[k]); idx = d.nonzero()[:3].tolist()
            print("run", it, "forward output", k, "differs:", int(d.sum()), "elements, first at", idx, "max abs diff",
                  float((cur[k].double() - base[k].double()).abs().max()), flush=True)
    for k in g:
        b = gbase[k].double(); tol = parity.RTOL * b.abs() + parity.GRAD_ATOL_REL * b.abs().max()
        v = float(((g[k].double() - b).abs() / tol).max()); worst = max(worst, v)
        if v > 0.5: print("run", it, "grad", k, "viol", v, flush=True)
print(f"stress {name}: {runs} runs, forward mismatches {bad_fwd}, worst grad viol {worst:.4f}")
sys.exit(1 if bad_fwd or worst > 0.5 else 0)
