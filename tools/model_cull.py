"""Work model: how many (tile, instance) and (8x4 block, instance) pairs survive an exact ellipse-vs-rectangle test
(minimum of the conic quadratic form over the rectangle against tau = 2 ln(255 op)) compared with the AABB test of the
alpha >= 1/255 extents that the kernels use today.  CPU, numpy, tile sample.  Guides kernel design; not product code."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scenegen, oracle

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 120
sc = scenegen.make_config(name)
cam = sc.cameras[0]
oracle.set_threads(3)
f = oracle.forward(sc, cam, render=False)
W, H = cam.image_width, cam.image_height
gx, gy = (W + 15) // 16, (H + 15) // 16
R = f["num_rendered"]
co = f["conic_opacity"]; m2 = f["means2D"]
A, B, Cc, op = co[:, 0].astype(np.float64), co[:, 1].astype(np.float64), co[:, 2].astype(np.float64), co[:, 3].astype(np.float64)
det = A * Cc - B * B
tau = 2.02 * np.log(np.maximum(255.0 * op, 1e-9)) + 0.02
with np.errstate(all="ignore"):
    ex = np.sqrt(tau * Cc / det) + 0.01
    ey = np.sqrt(tau * A / det) + 0.01
never = op < 1 / 255.0
ex[never] = -3e38; ey[never] = -3e38


def qmin_rect(a, b, c, cx, cy, x0, x1, y0, y1):
    """min over [x0,x1]x[y0,y1] of a dx^2 + 2 b dx dy + c dy^2, (dx,dy) = (cx-x, cy-y); convex -> interior or edges"""
    inside = (cx >= x0) & (cx <= x1) & (cy >= y0) & (cy <= y1)
    best = np.full(a.shape, np.inf)
    for xe in (x0, x1):  # vertical edges: dx fixed, minimise over y
        dx = cx - xe
        dy = np.clip(-b * dx / c, cy - y1, cy - y0)  # dy in [cy-y1, cy-y0]
        best = np.minimum(best, a * dx * dx + 2 * b * dx * dy + c * dy * dy)
    for ye in (y0, y1):
        dy = cy - ye
        dx = np.clip(-b * dy / a, cx - x1, cx - x0)
        best = np.minimum(best, a * dx * dx + 2 * b * dx * dy + c * dy * dy)
    return np.where(inside, 0.0, best)


rng = np.random.default_rng(0)
tiles = rng.choice(gx * gy, nsamp, replace=False)
tot = dict(inst=0, tile_aabb=0, tile_exact=0, blk_aabb=0, blk_exact=0, blk_pm=0, half_exact=0, half_pm=0, px_blend=0)
for t in tiles:
    ty, tx = divmod(t, gx)
    r0, r1 = f["ranges"][t]
    ids = f["point_list"][r0:r1]
    n = len(ids)
    tot["inst"] += n
    if n == 0: continue
    x, y = m2[ids, 0].astype(np.float64), m2[ids, 1].astype(np.float64)
    x0, y0 = tx * 16, ty * 16
    keep = (x + ex[ids] >= x0) & (x - ex[ids] <= x0 + 15) & (y + ey[ids] >= y0) & (y - ey[ids] <= y0 + 15)
    tot["tile_aabb"] += keep.sum()
    q = qmin_rect(A[ids], B[ids], Cc[ids], x, y, x0, x0 + 15, y0, y0 + 15)
    keep_e = keep & (q <= tau[ids])
    tot["tile_exact"] += keep_e.sum()
    ids = ids[keep]; x = x[keep]; y = y[keep]
    if len(ids) == 0: continue
    py, px = np.mgrid[y0:y0 + 16, x0:x0 + 16]
    dx = x[:, None, None] - px[None]; dy = y[:, None, None] - py[None]
    power = -0.5 * (A[ids, None, None] * dx * dx + Cc[ids, None, None] * dy * dy) - B[ids, None, None] * dx * dy
    alpha = np.minimum(0.99, op[ids, None, None] * np.exp(np.minimum(power, 0)))
    ok = (power <= 0) & (alpha >= 1 / 255.0)
    T = np.ones((16, 16)); done = np.zeros((16, 16), bool)
    blend = np.zeros_like(ok)
    alive = np.zeros((len(ids), 16, 16), bool)
    for k in range(len(ids)):
        alive[k] = ~done
        a = np.where(ok[k], alpha[k], 0.0)
        testT = T * (1 - a)
        stop = ok[k] & ~done & (testT < 1e-4)
        b = ok[k] & ~done & ~stop
        done |= stop
        T = np.where(b, testT, T)
        blend[k] = b
    tot["px_blend"] += blend.sum()
    for w in range(8):
        wy, wx = y0 + (w >> 1) * 4, x0 + (w & 1) * 8
        al = alive[:, (w >> 1) * 4:(w >> 1) * 4 + 4, (w & 1) * 8:(w & 1) * 8 + 8].reshape(len(ids), -1).any(1)
        hit = (x + ex[ids] >= wx) & (x - ex[ids] <= wx + 7) & (y + ey[ids] >= wy) & (y - ey[ids] <= wy + 3) & al
        tot["blk_aabb"] += hit.sum()
        q = qmin_rect(A[ids], B[ids], Cc[ids], x, y, wx, wx + 7, wy, wy + 3)
        tot["blk_exact"] += (hit & (q <= tau[ids])).sum()
        bw = blend[:, (w >> 1) * 4:(w >> 1) * 4 + 4, (w & 1) * 8:(w & 1) * 8 + 8]
        tot["blk_pm"] += bw.reshape(len(ids), -1).any(1).sum()
        for h in range(2):  # 8x2 half blocks
            hy = wy + 2 * h
            alh = alive[:, hy - y0:hy - y0 + 2, wx - x0:wx - x0 + 8].reshape(len(ids), -1).any(1)
            q = qmin_rect(A[ids], B[ids], Cc[ids], x, y, wx, wx + 7, hy, hy + 1)
            tot["half_exact"] += (alh & (q <= tau[ids])).sum()
            tot["half_pm"] += blend[:, hy - y0:hy - y0 + 2, wx - x0:wx - x0 + 8].reshape(len(ids), -1).any(1).sum()
scale = gx * gy / nsamp
print(name, "R", R, {k: f"{v * scale:.3e}" for k, v in tot.items()})
