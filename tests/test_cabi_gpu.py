"""GPU compute through the RAW C ABI (include/f3dgs_b200.h), without torch anywhere on the path: device memory from
cudaMalloc via ctypes, (function pointer, context) allocators, a non-default stream, results copied back with
cudaMemcpy and compared with the CPU oracle.  This is what a C-level embedder of the reference's inner interface
(`CudaRasterizer::Rasterizer`, rasterizer.h:18-94) would do -- see INTEGRATION.md section 3.
"""
import ctypes
import os

import numpy as np
import pytest

import parity
import scenegen

pytestmark = pytest.mark.gpu

ALLOC = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)


class Layout(ctypes.Structure):  # f3dgs_layout, include/f3dgs_b200.h
    _fields_ = [(n, ctypes.c_size_t) for n in (
        "geom_bytes", "geom_rec", "geom_cov3d", "geom_clamped", "geom_tiles", "geom_offsets", "geom_radii",
        "img_bytes", "img_final_T", "img_n_contrib", "img_ranges", "bin_bytes", "bin_point_list", "bin_keys")]

H2D, D2H = 1, 2


def _cudart():
    for name in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            return ctypes.CDLL(name)
        except OSError:
            pass
    import glob
    import site

    for sp in site.getsitepackages():
        for p in glob.glob(os.path.join(sp, "nvidia", "cuda_runtime", "lib", "libcudart.so*")):
            return ctypes.CDLL(p)
    pytest.skip("no libcudart.so found")


class Dev:
    """Tiny cudaMalloc arena."""

    def __init__(self, rt):
        self.rt, self.ptrs = rt, []

    def alloc(self, nbytes):
        p = ctypes.c_void_p()
        assert self.rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(max(int(nbytes), 256))) == 0
        self.ptrs.append(p)
        return p

    def put(self, a):
        a = np.ascontiguousarray(a)
        p = self.alloc(a.nbytes)
        assert self.rt.cudaMemcpy(p, a.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(a.nbytes), H2D) == 0
        return p

    def zeros(self, shape, dtype=np.float32):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.alloc(n)
        assert self.rt.cudaMemset(p, 0, ctypes.c_size_t(max(n, 1))) == 0
        return p

    def get(self, p, shape, dtype=np.float32):
        out = np.empty(shape, dtype)
        if out.nbytes:
            assert self.rt.cudaMemcpy(out.ctypes.data_as(ctypes.c_void_p), p, ctypes.c_size_t(out.nbytes), D2H) == 0
        return out

    def free(self):
        for p in self.ptrs:
            self.rt.cudaFree(p)


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_forward_backward_through_the_c_abi_without_torch(built, name):
    rt = _cudart()
    lib = ctypes.CDLL(built)
    lib.f3dgs_last_error.restype = ctypes.c_char_p
    lib.f3dgs_forward.restype = ctypes.c_int
    lib.f3dgs_backward.restype = ctypes.c_int
    sc = scenegen.make_config(name)
    cam = sc.cameras[0]
    P, C, W, H, M = sc.P, sc.C, cam.image_width, cam.image_height, sc.shs.shape[1]
    d = Dev(rt)
    stream = ctypes.c_void_p()
    assert rt.cudaStreamCreateWithFlags(ctypes.byref(stream), 1) == 0  # cudaStreamNonBlocking
    try:
        means, scales, rots = d.put(sc.means3D), d.put(sc.scales), d.put(sc.rotations)
        opac, shs, feats = d.put(sc.opacities), d.put(sc.shs), d.put(sc.features.reshape(P, C))
        bg, vm, pm, cp = d.put(sc.bg), d.put(cam.viewmatrix), d.put(cam.projmatrix), d.put(cam.campos)
        out_color, out_feat, out_depth = d.alloc(3 * H * W * 4), d.alloc(C * H * W * 4), d.alloc(H * W * 4)
        radii = d.alloc(P * 4)
        bufs = {}

        def mk(key):
            def cb(_ctx, nbytes):
                bufs[key] = d.alloc(nbytes)
                return bufs[key].value

            return ALLOC(cb)

        cbs = [mk("geom"), mk("bin"), mk("img")]
        null = ctypes.c_void_p(0)
        f = ctypes.c_float
        R = lib.f3dgs_forward(cbs[0], null, cbs[1], null, cbs[2], null, P, sc.sh_degree, M, C, bg, W, H, means, shs,
                              null, feats, opac, scales, f(1.0), rots, null, vm, pm, cp, f(cam.tanfovx),
                              f(cam.tanfovy), 0, out_color, out_feat, out_depth, radii, 0, stream)
        assert R >= 0, lib.f3dgs_last_error()
        assert rt.cudaStreamSynchronize(stream) == 0
        gc, gf, gd = scenegen.upstream_grads(H, W, C)
        dgc, dgf, dgd = d.put(gc), d.put(gf), d.put(gd)
        g = dict(mean2D=d.zeros((P, 3)), conic=d.zeros((P, 4)), opacity=d.zeros((P,)), color=d.zeros((P, 3)),
                 feat=d.zeros((P, C)), mean3D=d.zeros((P, 3)), cov3D=d.zeros((P, 6)), sh=d.zeros((P, M, 3)),
                 scale=d.zeros((P, 3)), rot=d.zeros((P, 4)), dz=d.zeros((P,)))
        rc = lib.f3dgs_backward(P, sc.sh_degree, M, R, C, bg, W, H, means, shs, null, feats, scales, f(1.0), rots,
                                null, vm, pm, cp, f(cam.tanfovx), f(cam.tanfovy), radii, bufs["geom"], bufs["bin"],
                                bufs["img"], dgc, dgf, dgd, g["mean2D"], g["conic"], g["opacity"], g["color"],
                                g["feat"], g["mean3D"], g["cov3D"], g["sh"], g["scale"], g["rot"], g["dz"], 0, stream)
        assert rc == 0, lib.f3dgs_last_error()
        assert rt.cudaStreamSynchronize(stream) == 0

        L = Layout()
        assert lib.f3dgs_get_layout(P, W, H, R, ctypes.byref(L)) == 0
        n_contrib = d.get(ctypes.c_void_p(bufs["img"].value + L.img_n_contrib), (H, W), np.uint32).astype(np.int64)
        point_list = d.get(ctypes.c_void_p(bufs["bin"].value + L.bin_point_list), (R,), np.uint32).astype(np.int64)
        ours = dict(num_rendered=np.int64(R), n_contrib=n_contrib, point_list=point_list,
                    color=d.get(out_color, (3, H, W)),
                    feature_map=d.get(out_feat, (C, H, W)), depth=d.get(out_depth, (1, H, W)),
                    radii=d.get(radii, (P,), np.int32))
        ours["grads"] = dict(means3D=d.get(g["mean3D"], (P, 3)), means2D=d.get(g["mean2D"], (P, 3)),
                             sh=d.get(g["sh"], (P, M, 3)), semantic_feature=d.get(g["feat"], (P, 1, C)),
                             opacities=d.get(g["opacity"], (P, 1)), scales=d.get(g["scale"], (P, 3)),
                             rotations=d.get(g["rot"], (P, 4)))
        orc = parity.run_oracle(sc, cam, grads=(gc, gf, gd), threads=1)
        rep = parity.compare(ours, orc, int_keys=("radii", "num_rendered", "point_list", "n_contrib"), float_keys=("color", "feature_map", "depth"),
                             tie_tolerant=True)
        print(parity.format_report(rep))
        assert rep["ok"], parity.format_report(rep)
    finally:
        rt.cudaStreamDestroy(stream)
        d.free()
