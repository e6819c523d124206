"""Properties of the CPU oracle that do not need the reference: they guard the checker itself
(culling, skip semantics, feature-width independence, and a finite-difference check of its backward)."""
import copy

import numpy as np
import pytest

import oracle
import parity
import scenegen


@pytest.fixture(scope="module")
def tiny():
    sc = scenegen.make_config("tiny")
    return sc, sc.cameras[0]


def test_behind_camera_is_culled_with_zero_gradients(tiny):
    sc, cam = tiny
    sc = copy.copy(sc)
    sc.means3D = sc.means3D.copy()
    sc.means3D[:50] = cam.campos * 2.0  # behind the camera (it looks at the origin)
    grads = scenegen.upstream_grads(cam.image_height, cam.image_width, sc.C)
    o = parity.run_oracle(sc, cam, grads=grads, threads=1)
    assert (o["radii"][:50] == 0).all()
    for k in ("means3D", "scales", "rotations", "opacities", "sh", "semantic_feature", "means2D"):
        assert np.abs(o["grads"][k][:50]).max() == 0, k


def test_zero_and_subthreshold_opacity_contribute_nothing(tiny):
    sc, cam = tiny
    base = parity.run_oracle(sc, cam, threads=1)
    sc2 = copy.copy(sc)
    sc2.opacities = sc.opacities.copy()
    drop = np.arange(sc.P) % 3 == 0
    sc2.opacities[drop] = 1.0 / 300.0
    a = parity.run_oracle(sc2, cam, threads=1)
    sc3 = copy.copy(sc)  # same cloud with those Gaussians removed altogether
    keep = ~drop
    for f in ("means3D", "scales", "rotations", "opacities", "shs", "features"):
        setattr(sc3, f, getattr(sc, f)[keep])
    b = parity.run_oracle(sc3, cam, threads=1)
    for k in ("color", "feature_map", "depth", "final_T"):
        assert np.array_equal(a[k], b[k]), k
    assert not np.array_equal(a["color"], base["color"])


def test_colour_depth_and_indices_do_not_depend_on_feature_width(tiny):
    sc, cam = tiny
    a = parity.run_oracle(sc, cam, threads=1)
    sc0 = copy.copy(sc)
    sc0.features = np.zeros((sc.P, 1, 0), np.float32)
    b = parity.run_oracle(sc0, cam, threads=1)
    for k in ("color", "depth", "final_T", "n_contrib", "point_list", "ranges", "radii"):
        assert np.array_equal(a[k], b[k]), k


def test_n_contrib_is_index_of_last_blended_and_T_matches(tiny):
    sc, cam = tiny
    o = parity.run_oracle(sc, cam, threads=1)
    f = o["fwd"]
    W = cam.image_width
    gx = (W + 15) // 16
    for (py, px) in [(3, 5), (20, 40), (55, 79), (31, 17)]:
        r0, r1 = f["ranges"][(py // 16) * gx + px // 16]
        T, last = 1.0, 0
        for i in range(int(r0), int(r1)):
            g = f["point_list"][i]
            dx, dy = f["means2D"][g] - np.array([px, py], np.float32)
            a, b, c, op = f["conic_opacity"][g]
            power = -0.5 * (a * dx * dx + c * dy * dy) - b * dx * dy
            if power > 0:
                continue
            alpha = min(0.99, op * np.exp(power))
            if alpha < 1 / 255:
                continue
            if T * (1 - alpha) < 1e-4:
                break
            T *= 1 - alpha
            last = i - int(r0) + 1
        assert abs(T - o["final_T"][py, px]) < 1e-5
        assert last == o["n_contrib"][py, px]


def test_backward_against_finite_differences():
    """Central differences on a 40-Gaussian scene for the loss L = <color, gc> + <depth, gd> (+ <feat, gf> for the
    feature input).  Reference quirks kept: the feature loss does not reach geometry (backward.cu:575 disabled), and
    the rotation gradient omits the normalisation Jacobian, so those are checked on the matching sub-losses only."""
    sc = scenegen.make_scene(P=40, W=48, H=32, C=4, sh_degree=2, seed=77, target_radius_px=7.0)
    cam = sc.cameras[0]
    sc.opacities[:] = np.clip(sc.opacities, 0.3, 0.9)
    gc, gf, gd = scenegen.upstream_grads(32, 48, 4, seed=5)

    def loss(s, with_feat):
        f = oracle.forward(s, cam)
        v = float((f["color"].astype(np.float64) * gc).sum() + (f["depth"].astype(np.float64) * gd).sum())
        if with_feat:
            v += float((f["feature_map"].astype(np.float64) * gf).sum())
        return v

    oracle.set_threads(1)
    f0 = oracle.forward(sc, cam)
    g_geo = oracle.backward(sc, cam, f0, gc, np.zeros_like(gf), gd)
    g_all = oracle.backward(sc, cam, f0, gc, gf, gd)
    rng = np.random.Generator(np.random.PCG64(3))

    def fd(field, idx, eps, with_feat):
        sp, sm = copy.copy(sc), copy.copy(sc)
        ap, am = getattr(sc, field).copy(), getattr(sc, field).copy()
        ap[idx] += eps
        am[idx] -= eps
        setattr(sp, field, ap)
        setattr(sm, field, am)
        return (loss(sp, with_feat) - loss(sm, with_feat)) / (2 * eps)

    checks = [("means3D", "means3D", 2e-4, g_geo), ("scales", "scales", 1e-5, g_geo),
              ("opacities", "opacities", 1e-3, g_geo), ("shs", "sh", 1e-2, g_geo)]
    for field, gname, eps, g in checks:
        arr = getattr(sc, field)
        num, ana = [], []
        for _ in range(12):
            idx = tuple(rng.integers(0, s) for s in arr.shape)
            if f0["radii"][idx[0]] == 0:
                continue
            num.append(fd(field, idx, eps, False))
            ana.append(float(g[gname][idx]))
        num, ana = np.array(num), np.array(ana)
        scale = max(np.abs(ana).max(), 1e-3)
        assert np.abs(num - ana).max() <= 0.03 * scale + 2e-3, (field, num, ana)
    # features: exact linear dependence
    num, ana = [], []
    for _ in range(10):
        idx = (int(rng.integers(0, sc.P)), 0, int(rng.integers(0, 4)))
        num.append(fd("features", idx, 1e-2, True))
        ana.append(float(g_all["semantic_feature"][idx]))
    assert np.allclose(num, ana, rtol=2e-3, atol=2e-4)
    # the feature loss moves nothing but the features (reference quirk D.1)
    for k in ("means3D", "scales", "rotations", "opacities", "sh"):
        assert np.array_equal(g_geo[k], g_all[k]), k
