"""GPU debugging aid for the two-pass mode (F3DGS_SPLIT=1|2): pull the per-(tile, block) instance lists out of the binning
buffer of one forward call and check them against a float32 numpy recomputation from the library's own per-Gaussian
records:  entry order and Gaussian ids, pixel masks, blend weights, and that  sum_entries w * feature  reproduces the
feature map the feature pass wrote.

    F3DGS_SPLIT=1 python tools/check_lists.py small [n_tiles]

Layout restated from csrc/api.cu (BinLayout, 256-byte aligned fields): point_list u32[R], keys u64[R], unsorted twins,
then list_w f32[8R][32], list_meta {id, mask}[8R], list_cnt u32[8T]; block b of tile t owns entries
[8*range.x + b*len, +len), len = range.y - range.x.  Lane l of a block is pixel (x, y) = (((l>>2)&3)*2 + (l&1),
(l>>4)*2 + ((l>>1)&1)) (composite_common.cuh: lane_px / lane_py).  Development tool, not product code."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "feature-3dgs_b200"))
assert os.environ.get("F3DGS_SPLIT") in ("1", "2"), "run with F3DGS_SPLIT=1 (or 2)"
import numpy as np  # noqa: E402
import torch  # noqa: E402

import scenegen  # noqa: E402
from diff_gaussian_rasterization import _C  # noqa: E402

f32 = np.float32


def align(x, a=256):
    return (x + a - 1) // a * a


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "small"
    nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    sc = scenegen.make_config(name)
    cam = sc.cameras[0]
    dev = "cuda"
    t = scenegen.to_torch(sc, dev, requires_grad=False)
    kw = scenegen.settings_kwargs(sc, cam, dev)
    P, C, H, W = sc.P, sc.C, cam.image_height, cam.image_width
    assert C > 0, "lists exist only for C > 0"
    empty = torch.Tensor([])
    R, color, fmap, depth, radii, geom, binb, img = _C.rasterize_gaussians(
        kw["bg"], t["means3D"], empty, t["semantic_feature"], t["opacities"], t["scales"], t["rotations"], 1.0, empty,
        kw["viewmatrix"], kw["projmatrix"], kw["tanfovx"], kw["tanfovy"], H, W, t["shs"], sc.sh_degree, kw["campos"],
        False, False)
    torch.cuda.synchronize()
    point_list, ranges, n_contrib, final_T, rec = [x.cpu().numpy() for x in _C.debug_views(geom, binb, img, P, W, H, R)]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    o = align(R * 4); o = align(o + R * 8); o = align(o + R * 4); o = align(o + R * 8)
    off_w = o; o = align(o + 8 * R * 128)
    off_m = o; o = align(o + 8 * R * 8)
    off_c = o
    raw = binb.cpu().numpy()
    assert raw.size >= off_c + 8 * T * 4, "binning buffer has no list region: was the library built with the two-pass mode?"
    list_w = raw[off_w:off_w + 8 * R * 128].view(np.float32).reshape(8 * R, 32)
    list_meta = raw[off_m:off_m + 8 * R * 8].view(np.uint32).reshape(8 * R, 2)
    list_cnt = raw[off_c:off_c + 8 * T * 4].view(np.uint32)
    feats = t["semantic_feature"].reshape(P, C).cpu().numpy()
    fmap = fmap.cpu().numpy()
    lane = np.arange(32)
    lx = ((lane >> 2) & 3) * 2 + (lane & 1)
    ly = (lane >> 4) * 2 + ((lane >> 1) & 1)

    rng = np.random.default_rng(0)
    tiles = rng.choice(T, min(nsamp, T), replace=False)
    bad = dict(count=0, gid=0, mask=0, w=0, fmap=0)
    worst_w = worst_f = 0.0
    entries = 0
    for tile in tiles:
        ty, tx = divmod(int(tile), gx)
        r0, r1 = int(ranges[tile, 0]), int(ranges[tile, 1])
        ids = point_list[r0:r1]
        L = r1 - r0
        for b in range(8):
            bx0, by0 = tx * 16 + (b & 1) * 8, ty * 16 + (b >> 1) * 4
            px, py = (bx0 + lx).astype(f32), (by0 + ly).astype(f32)
            inside = (bx0 + lx < W) & (by0 + ly < H)
            Tcur = np.ones(32, f32)
            done = ~inside
            exp_gid, exp_pm, exp_w = [], [], []
            for g in ids:
                x, y, a, bb, c, op = (f32(rec[g, 0]), f32(rec[g, 1]), f32(rec[g, 4]), f32(rec[g, 5]), f32(rec[g, 6]),
                                      f32(rec[g, 7]))
                dx, dy = x - px, y - py
                power = (f32(-0.5) * (a * dx * dx + c * dy * dy) - bb * dx * dy).astype(f32)
                with np.errstate(all="ignore"):
                    alpha = np.minimum(f32(0.99), op * np.exp(power).astype(f32)).astype(f32)
                ok = (power <= 0) & (alpha >= f32(1 / 255.0)) & ~done
                testT = (Tcur * (f32(1) - alpha)).astype(f32)
                stop = ok & (testT < f32(1e-4))
                blend = ok & ~stop
                done = done | stop
                if blend.any():
                    exp_gid.append(int(g))
                    exp_pm.append(int(np.sum((1 << lane)[blend])))
                    exp_w.append(np.where(blend, alpha * Tcur, f32(0)).astype(f32))
                Tcur = np.where(blend, testT, Tcur)
            base = 8 * r0 + b * L
            n = int(list_cnt[tile * 8 + b])
            entries += n
            if n != len(exp_gid):
                bad["count"] += 1
                continue  # threshold flips (alpha ~ 1/255, T ~ 1e-4) can legitimately add or drop an entry: counted, not fatal
            gid = list_meta[base:base + n, 0]
            pm = list_meta[base:base + n, 1]
            w = list_w[base:base + n]
            bad["gid"] += int((gid != np.array(exp_gid, np.uint32)).sum())
            bad["mask"] += int((pm != np.array(exp_pm, np.uint32)).sum())
            if n:
                ew = np.stack(exp_w)
                err = np.abs(w - ew).max() / max(float(np.abs(ew).max()), 1e-20)
                worst_w = max(worst_w, float(err))
                bad["w"] += int(err > 1e-4)
                # feature map of the block from the lists the library wrote (float64 accumulate)
                acc = np.einsum("np,nc->pc", w.astype(np.float64), feats[gid].astype(np.float64))
            else:
                acc = np.zeros((32, C))
            got = np.stack([fmap[:, by0 + ly[l], bx0 + lx[l]] if inside[l] else np.zeros(C) for l in range(32)])
            scale = max(float(np.abs(got).max()), 1e-12)
            ef = float(np.abs(np.where(inside[:, None], got - acc, 0)).max()) / scale
            worst_f = max(worst_f, ef)
            bad["fmap"] += int(ef > 1e-4)
    print(f"{name}: {len(tiles)} tiles, {entries} entries checked; mismatches {bad}; worst weight error {worst_w:.2e}, "
          f"worst feature-map error {worst_f:.2e} (relative to the block maximum)")
    ok = bad["gid"] == 0 and bad["mask"] <= 2 and bad["w"] == 0 and bad["fmap"] == 0 and bad["count"] <= 2
    print("LISTS OK" if ok else "LISTS MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
