"""GPU bring-up: parity of this repo's CUDA path vs the reference extension (oracle/_ref) and the CPU oracle,
plus a first timing.  Usage: python tools/gpu_bringup.py <step> ...   steps: tiny small c1 c2 c3 time_c2 time_c3
Writes gpurun_out/bringup_<step>.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
import scenegen  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def parity_step(name, with_oracle, debug):
    import torch
    from oracle import ref_wrapper as rw

    sc = scenegen.make_config(name)
    cam = sc.cameras[0]
    grads = scenegen.upstream_grads(cam.image_height, cam.image_width, sc.C)
    res = {}
    t = time.time()
    ours = parity.run_ours(sc, cam, grads=grads, debug=debug)
    torch.cuda.synchronize()
    print(f"[{name}] ours done in {time.time() - t:.2f}s  R={int(ours['num_rendered'])}", flush=True)
    if rw.available(sc.C):
        ref = parity.run_ref(sc, cam, grads=grads)
        rep = parity.compare(ours, ref)
        print(f"[{name}] ours vs reference CUDA:\n" + parity.format_report(rep), flush=True)
        res["vs_ref"] = rep
        # per-Gaussian intermediates
        vis = ref["radii"] > 0
        for k, sl in (("means2D", slice(0, 2)), ("conic_opacity", slice(4, 8)), ("rgb", slice(8, 11))):
            a = ours["rec"][vis][:, sl]
            b = ref["geom"][k][vis]
            print(f"    rec.{k}: bit-exact={np.array_equal(a, b)} maxdiff={np.abs(a - b).max() if a.size else 0}")
        a, b = ours["rec"][vis][:, 11], ref["geom"]["depths"][vis]
        print(f"    rec.depth: bit-exact={np.array_equal(a, b)}")
        # reference run-to-run spread of its own gradients (atomics)
        ref2 = parity.run_ref(sc, cam, grads=grads)
        spread = {k: parity.float_mismatch(ref2["grads"][k], ref["grads"][k])[0] for k in ref["grads"]}
        print(f"[{name}] reference self-spread (viol ratio): " + ", ".join(f"{k}={v:.2g}" for k, v in spread.items()))
        res["ref_self_spread"] = spread
    if with_oracle:
        orc = parity.run_oracle(sc, cam, grads=grads, threads=1)
        rep = parity.compare(ours, orc)
        print(f"[{name}] ours vs CPU oracle:\n" + parity.format_report(rep), flush=True)
        res["vs_oracle"] = rep
    with open(os.path.join(OUT, f"bringup_{name}.json"), "w") as f:
        json.dump(res, f, indent=1, default=float)


def time_step(name, iters=5):
    import torch
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from oracle import ref_wrapper as rw

    sc = scenegen.make_config(name)
    cam = sc.cameras[0]
    dev = "cuda"
    grads = [torch.from_numpy(g).to(dev) for g in scenegen.upstream_grads(cam.image_height, cam.image_width, sc.C)]
    res = {}
    for impl in ("ours", "ref"):
        if impl == "ref" and not rw.available(sc.C):
            continue
        t = scenegen.to_torch(sc, dev, requires_grad=True)
        rsk = scenegen.settings_kwargs(sc, cam, dev)
        if impl == "ours":
            rast = GaussianRasterizer(GaussianRasterizationSettings(**rsk))
        else:
            rast = rw.RefRasterizer(rsk, sc.C)
        fw, bw = [], []
        for it in range(iters + 2):
            means2D = torch.zeros_like(t["means3D"], requires_grad=True)
            for k in t:
                t[k].grad = None
            e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e0.record()
            color, feat, radii, depth = rast(means3D=t["means3D"], means2D=means2D, opacities=t["opacities"],
                                             shs=t["shs"], semantic_feature=t["semantic_feature"] if sc.C else None,
                                             scales=t["scales"], rotations=t["rotations"])
            e1.record()
            outs, gos = [color, depth], [grads[0], grads[2]]
            if sc.C:
                outs.append(feat)
                gos.append(grads[1])
            torch.autograd.backward(outs, gos)
            e2.record()
            torch.cuda.synchronize()
            if it >= 2:
                fw.append(e0.elapsed_time(e1))
                bw.append(e1.elapsed_time(e2))
        res[impl] = dict(fwd_ms=float(np.median(fw)), bwd_ms=float(np.median(bw)))
        print(f"[time {name}] {impl}: fwd {np.median(fw):.3f} ms  bwd {np.median(bw):.3f} ms  "
              f"-> {1000.0 / (np.median(fw) + np.median(bw)):.1f} views/s", flush=True)
        del t, rast
        torch.cuda.empty_cache()
    with open(os.path.join(OUT, f"bringup_time_{name}.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    for step in sys.argv[1:]:
        if step.startswith("time_"):
            time_step(step[5:])
        else:
            parity_step(step, with_oracle=step in ("tiny", "small", "c1"), debug=step == "tiny")
