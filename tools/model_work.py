"""Work model of the composite kernels on a tile sample (CPU, numpy): how many instances survive each culling
level and how dense the blend-weight matrix is.  Guides kernel design; not part of the product."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scenegen, oracle

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 120
sc = scenegen.make_config(name)
cam = sc.cameras[0]
oracle.set_threads(8)
f = oracle.forward(sc, cam, render=False)
W, H = cam.image_width, cam.image_height
gx, gy = (W + 15) // 16, (H + 15) // 16
R = f["num_rendered"]
print("P", sc.P, "R", R, "tiles", gx * gy, "inst/tile", R / (gx * gy))
co = f["conic_opacity"]; m2 = f["means2D"]
A, B, Cc, op = co[:, 0], co[:, 1], co[:, 2], co[:, 3]
det = A * Cc - B * B
tau = 2.02 * np.log(np.maximum(255.0 * op, 1e-9)) + 0.02
with np.errstate(all="ignore"):
    ex = np.sqrt(tau * Cc / det) + 0.01
    ey = np.sqrt(tau * A / det) + 0.01
never = op < 1 / 255.0
ex[never] = -3e38; ey[never] = -3e38
rng = np.random.default_rng(0)
tiles = rng.choice(gx * gy, nsamp, replace=False)
tot = dict(inst=0, tile_keep=0, warp_hit=0, warp_pm=0, px_blend=0, quad2x2=0, quad4x1=0, pair_eval=0, stages=0, px_alpha_ok=0)
wl = []
for t in tiles:
    ty, tx = divmod(t, gx)
    r0, r1 = f["ranges"][t]
    ids = f["point_list"][r0:r1]
    n = len(ids)
    tot["inst"] += n
    if n == 0: continue
    x, y = m2[ids, 0], m2[ids, 1]
    x0, y0 = tx * 16, ty * 16
    keep = (x + ex[ids] >= x0) & (x - ex[ids] <= x0 + 15) & (y + ey[ids] >= y0) & (y - ey[ids] <= y0 + 15)
    ids = ids[keep]; x = x[keep]; y = y[keep]
    tot["tile_keep"] += len(ids)
    tot["stages"] += (len(ids) + 31) // 32 + 1
    if len(ids) == 0: continue
    py, px = np.mgrid[y0:y0 + 16, x0:x0 + 16]
    dx = x[:, None, None] - px[None]; dy = y[:, None, None] - py[None]
    power = -0.5 * (A[ids, None, None] * dx * dx + Cc[ids, None, None] * dy * dy) - B[ids, None, None] * dx * dy
    alpha = np.minimum(0.99, op[ids, None, None] * np.exp(np.minimum(power, 0)))
    ok = (power <= 0) & (alpha >= 1 / 255.0)
    tot["px_alpha_ok"] += ok.sum()
    # sequential T
    T = np.ones((16, 16)); done = np.zeros((16, 16), bool)
    blend = np.zeros_like(ok)
    alive_warp = np.zeros((len(ids), 8), bool)
    for k in range(len(ids)):
        a = np.where(ok[k], alpha[k], 0.0)
        testT = T * (1 - a)
        stop = ok[k] & ~done & (testT < 1e-4)
        b = ok[k] & ~done & ~stop
        done |= stop
        T = np.where(b, testT, T)
        blend[k] = b
        for w in range(8):
            wy, wx = (w >> 1) * 4, (w & 1) * 8
            alive_warp[k, w] = (~done[wy:wy + 4, wx:wx + 8]).any()
    tot["px_blend"] += blend.sum()
    for w in range(8):
        wy, wx = y0 + (w >> 1) * 4, x0 + (w & 1) * 8
        hit = (x + ex[ids] >= wx) & (x - ex[ids] <= wx + 7) & (y + ey[ids] >= wy) & (y - ey[ids] <= wy + 3) & alive_warp[:, w]
        tot["warp_hit"] += hit.sum()
        bw = blend[:, (w >> 1) * 4:(w >> 1) * 4 + 4, (w & 1) * 8:(w & 1) * 8 + 8]
        pm = bw.reshape(len(ids), -1).any(1)
        tot["warp_pm"] += pm.sum()
        q22 = bw.reshape(len(ids), 2, 2, 4, 2).any(axis=(2, 4)).sum()
        q41 = bw.reshape(len(ids), 4, 2, 4).any(axis=3).sum()
        tot["quad2x2"] += q22; tot["quad4x1"] += q41
scale = gx * gy / nsamp
print({k: f"{v * scale:.3e}" for k, v in tot.items()})
print("per-pixel blended:", tot["px_blend"] / (nsamp * 256), " alpha_ok per px:", tot["px_alpha_ok"] / (nsamp * 256))
print("tile keep frac", tot["tile_keep"] / tot["inst"], "warp hit per kept", tot["warp_hit"] / (8 * tot["tile_keep"]),
      "pm per hit", tot["warp_pm"] / tot["warp_hit"], "px per pm", tot["px_blend"] / tot["warp_pm"],
      "quads2x2 per pm", tot["quad2x2"] / tot["warp_pm"], "4x1 per pm", tot["quad4x1"] / tot["warp_pm"])
