"""Multi-rank host logic on CPU (gloo, world_size 2): view sharding + ONE flat-buffer gradient all-reduce per
step must reproduce the single-process sum of per-view gradients.  The per-view render here is the CPU oracle
wrapped in an autograd.Function (test infrastructure -- the product path has no CPU renderer)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _paths():
    for p in (ROOT, os.path.join(ROOT, "feature-3dgs_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)


class _OracleRender(torch.autograd.Function):
    """CPU autograd bridge over the oracle (means3D, scales, rotations, opacities, shs, features) -> images."""

    @staticmethod
    def forward(ctx, means3D, scales, rotations, opacities, shs, features, scene, cam):
        import copy

        import oracle

        sc = copy.copy(scene)
        sc.means3D, sc.scales, sc.rotations = means3D.numpy(), scales.numpy(), rotations.numpy()
        sc.opacities, sc.shs, sc.features = opacities.numpy(), shs.numpy(), features.numpy()
        f = oracle.forward(sc, cam)
        ctx.sc, ctx.cam, ctx.f = sc, cam, f
        return (torch.from_numpy(f["color"].copy()), torch.from_numpy(f["feature_map"].copy()),
                torch.from_numpy(f["depth"].copy()))

    @staticmethod
    def backward(ctx, gc, gf, gd):
        import oracle

        g = oracle.backward(ctx.sc, ctx.cam, ctx.f, gc.numpy(), gf.numpy(), gd.numpy())
        t = torch.from_numpy
        return (t(g["means3D"]), t(g["scales"]), t(g["rotations"]), t(g["opacities"]), t(g["sh"]),
                t(g["semantic_feature"]), None, None)


def _step(scene, view_ids, params, grads_up):
    gc, gf, gd = grads_up
    losses = []
    for v in view_ids:
        color, feat, depth = _OracleRender.apply(*params, scene, scene.cameras[v])
        loss = (color * gc).sum() + (feat * gf).sum() + (depth * gd).sum()
        loss.backward()
        losses.append(float(loss.detach()))
    return losses


def _make(n_views):
    _paths()
    import oracle
    import scenegen

    oracle.set_threads(1)
    sc = scenegen.make_scene(P=300, W=48, H=32, C=4, sh_degree=1, views=n_views, seed=11)
    t = scenegen.to_torch(sc, "cpu", requires_grad=True)
    params = [t[k] for k in ("means3D", "scales", "rotations", "opacities", "shs", "semantic_feature")]
    ups = [torch.from_numpy(g) for g in scenegen.upstream_grads(32, 48, 4, seed=3)]
    return sc, params, ups


def _worker(rank, world, port, n_views, out_path):
    _paths()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diff_gaussian_rasterization.parallel import FlatGradBuffer, shard_views

    sc, params, ups = _make(n_views)
    flat = FlatGradBuffer(params, extra=2)
    mine = shard_views(n_views, rank, world)
    _step(sc, mine, params, ups)
    flat.extra[0] = float(len(mine))  # a step statistic riding in the same collective
    flat.check_views()
    flat.all_reduce()
    if rank == 0:
        np.save(out_path, flat.flat.numpy())
    dist.destroy_process_group()


def test_shard_views_partitions_every_view_once():
    _paths()
    from diff_gaussian_rasterization.parallel import shard_views

    for n, w in [(64, 8), (7, 2), (3, 4), (0, 2)]:
        got = sorted(v for r in range(w) for v in shard_views(n, r, w))
        assert got == list(range(n))
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)


def test_flat_grad_buffer_aliases_param_grads():
    _paths()
    from diff_gaussian_rasterization.parallel import FlatGradBuffer

    a = torch.zeros(5, 3, requires_grad=True)
    b = torch.zeros(7, requires_grad=True)
    flat = FlatGradBuffer([a, b])
    (a.sum() * 2 + (b * torch.arange(7.0)).sum()).backward()
    (a.sum()).backward()  # second "view": accumulates in place
    flat.check_views()
    assert torch.equal(flat.flat[:15], torch.full((15,), 3.0)) and torch.equal(flat.flat[15:], torch.arange(7.0))
    flat.zero_()
    assert float(a.grad.abs().sum()) == 0
    with pytest.raises(ValueError):
        FlatGradBuffer([torch.zeros(3)])


@pytest.mark.timeout(300)
def test_two_rank_all_reduce_equals_single_process_sum(tmp_path):
    n_views, world = 4, 2
    out = str(tmp_path / "flat.npy")
    mp.spawn(_worker, args=(world, _free_port(), n_views, out), nprocs=world, join=True)
    got = np.load(out)
    # single process, all views
    _paths()
    from diff_gaussian_rasterization.parallel import FlatGradBuffer

    sc, params, ups = _make(n_views)
    flat = FlatGradBuffer(params, extra=2)
    _step(sc, range(n_views), params, ups)
    want = flat.flat.numpy().copy()
    want[-2] = n_views
    assert got.shape == want.shape
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-5 * scale  # summation order differs (2+2 vs 4 sequential)
    assert got[-2] == n_views
