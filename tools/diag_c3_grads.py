"""Where do ours and the reference CUDA build differ in the c3 gradients, and how does that compare with the
reference's own run-to-run spread (its atomicAdd order is not deterministic)?  usage: diag_c3_grads.py [cfg]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "feature-3dgs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenegen, parity
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
sc = scenegen.make_config(name); cam = sc.cameras[0]
grads = scenegen.upstream_grads(cam.image_height, cam.image_width, sc.C)
o1 = parity.run_ours(sc, cam, grads=grads)["grads"]
o2 = parity.run_ours(sc, cam, grads=grads)["grads"]
r1 = parity.run_ref(sc, cam, grads=grads)["grads"]
r2 = parity.run_ref(sc, cam, grads=grads)["grads"]
def viol(a, b, atol=parity.GRAD_ATOL_REL):
    a = a.astype(np.float64); b = b.astype(np.float64)
    tol = parity.RTOL * np.abs(b) + atol * np.abs(b).max()
    v = np.abs(a - b) / tol
    i = np.unravel_index(np.argmax(v), v.shape)
    return float(v.max()), i, float(a[i]), float(b[i]), float(np.abs(b).max()), int((v > 1).sum())
for k in parity.GRAD_KEYS:
    vo, i, a, b, mx, nbad = viol(o1[k], r1[k])
    print(f"{k:18s} ours-ref viol {vo:8.3f} at {i} ours {a:+.6e} ref {b:+.6e} max|ref| {mx:.3e} nbad {nbad} | "
          f"ref-ref2 {viol(r2[k], r1[k])[0]:8.3f} ours-ours2 {viol(o2[k], o1[k])[0]:8.3f}", flush=True)
    if vo > 1:
        g = i[0]
        print("   gaussian", g, "ours", o1[k][g].ravel()[:4], "ref", r1[k][g].ravel()[:4], "ref2", r2[k][g].ravel()[:4], "ours2", o2[k][g].ravel()[:4])
        for kk in ("means2D", "opacities", "scales"):
            print("     ", kk, "ours", o1[kk][g].ravel()[:3], "ref", r1[kk][g].ravel()[:3])
        print("      mean", sc.means3D[g], "scale", sc.scales[g], "opacity", sc.opacities[g])
