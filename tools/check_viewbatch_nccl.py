"""world_size-N NCCL check of ViewBatch (run under torchrun on a multi-GPU box): the all-reduced flat gradient buffer of N
ranks rendering N*V sharded views equals the buffer of one rank rendering all of them (two-bucket overlapped all-reduce
included).  Prints OK / FAIL per rank-0.  usage: torchrun --nproc-per-node N tools/check_viewbatch_nccl.py [config] [views/rank]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "feature-3dgs_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import scenegen  # noqa: E402
from diff_gaussian_rasterization import GaussianRasterizationSettings  # noqa: E402
from diff_gaussian_rasterization.parallel import ViewBatch, shard_views  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "small128"
vpr = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
sc = scenegen.make_config(cfg, views=vpr * world)
t = scenegen.to_torch(sc, dev)
cam0 = sc.cameras[0]
ups = [torch.from_numpy(g).to(dev) for g in scenegen.upstream_grads(cam0.image_height, cam0.image_width, sc.C)]


def run(views, reduce):
    vb = ViewBatch({k: t[k] for k in ("means3D", "scales", "rotations", "opacities", "shs", "semantic_feature")})
    vb.zero_()
    for i, v in enumerate(views):
        rs = GaussianRasterizationSettings(**scenegen.settings_kwargs(sc, sc.cameras[v], dev))
        color, feat, radii, depth, ctx = vb.forward(rs)
        vb.backward(ctx, ups[0], ups[1], ups[2], last=(i == len(views) - 1))
    if reduce:
        vb.all_reduce()
    torch.cuda.synchronize()
    return vb.flat.clone()


mine = run(shard_views(vpr * world, rank, world), True)
if rank == 0:
    full = run(list(range(vpr * world)), False)
    err = (mine - full).abs().max().item()
    scale = full.abs().max().item()
    ok = err <= 5e-5 * scale
    print(f"viewbatch nccl world={world} cfg={cfg}: max |diff| {err:.3e} of scale {scale:.3e} -> {'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        sys.exit(1)
dist.barrier(device_ids=[lr])
dist.destroy_process_group()
