// Post-raster feature head: bilinear resize of the rendered feature map to the teacher's
// resolution fused with the L1 feature loss and its gradient.
// Reference: train.py:98-104
//     feature_map = F.interpolate(feature_map.unsqueeze(0), size=gt.shape[1:], mode='bilinear', align_corners=True)
//     [feature_map = cnn_decoder(feature_map)]                       (models/networks.py:107-119, only with --speedup)
//     Ll1_feature = l1_loss(feature_map, gt_feature_map)             (utils/loss_utils.py: mean |a - b|)
// In PyTorch that is a resize kernel (write C*Hg*Wg), an L1 kernel pair, and in the backward a sign kernel, a scatter of
// the resize gradient with atomics into a zero-filled C*H*W tensor.  Here:
//     resize_fwd   reads the C*H*W map once; with a target it writes  sign(interp - gt) * grad_scale  (the gradient of the
//                  loss w.r.t. the resized map, ready for the backward) and accumulates sum |interp - gt|; without a
//                  target it writes the resized map (decoder path: the 1x1 convolution in between stays a library GEMM).
//     resize_bwd   GATHERS: one thread per element of the C*H*W gradient sums the (few) target pixels whose bilinear
//                  footprint covers it -- every element written exactly once, no atomics, no zero fill.
// Sampling positions follow ATen's upsample_bilinear2d with align_corners=True: src = dst * (in - 1) / (out - 1),
// i0 = (int)src, i1 = i0 + (i0 < in - 1), lambda1 = src - i0.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/f3dgs_b200.h"
#include "kernels.h"

namespace f3dgs {
namespace {

struct ResizeGeom {
    int C, H, W, Hg, Wg;
    float ry, rx;  // (in - 1) / (out - 1), 0 when out == 1
};

__device__ __forceinline__ void src_of(int o, float r, int in, int& i0, int& i1, float& l0, float& l1) {
    const float s = r * (float)o;
    i0 = (int)s;
    i1 = i0 + ((i0 < in - 1) ? 1 : 0);
    l1 = s - (float)i0;
    l0 = 1.0f - l1;
}

// one thread per element of the [C, Hg, Wg] output, x fastest
__global__ void __launch_bounds__(256) resize_fwd_kernel(ResizeGeom g, const float* __restrict__ fm,
                                                         const float* __restrict__ gt, float grad_scale,
                                                         float* __restrict__ out, float* __restrict__ loss_sum) {
    const size_t n = (size_t)g.C * g.Hg * g.Wg;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float ad = 0.f;
    if (i < n) {
        const int ox = (int)(i % g.Wg);
        const int oy = (int)((i / g.Wg) % g.Hg);
        const int c = (int)(i / ((size_t)g.Wg * g.Hg));
        int y0, y1, x0, x1;
        float ly0, ly1, lx0, lx1;
        src_of(oy, g.ry, g.H, y0, y1, ly0, ly1);
        src_of(ox, g.rx, g.W, x0, x1, lx0, lx1);
        const float* p = fm + (size_t)c * g.H * g.W;
        const float v = ly0 * (lx0 * __ldg(p + (size_t)y0 * g.W + x0) + lx1 * __ldg(p + (size_t)y0 * g.W + x1)) +
                        ly1 * (lx0 * __ldg(p + (size_t)y1 * g.W + x0) + lx1 * __ldg(p + (size_t)y1 * g.W + x1));
        if (gt != nullptr) {
            const float d = v - gt[i];
            ad = fabsf(d);
            out[i] = d > 0.f ? grad_scale : (d < 0.f ? -grad_scale : 0.f);
        } else {
            out[i] = v;
        }
    }
    if (gt != nullptr && loss_sum != nullptr) {
        // block reduction of |d|, one atomic per block
        __shared__ float part[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ad += __shfl_xor_sync(0xffffffffu, ad, o);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ad;
        __syncthreads();
        if (threadIdx.x < 8) {
            float s = part[threadIdx.x];
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
            if (threadIdx.x == 0) atomicAdd(loss_sum, s);
        }
    }
}

// Output rows whose footprint can touch source row y: src = r * o in (y - 1, y + 1)  ->  a conservative integer range,
// each candidate is then tested with the forward's own arithmetic.
__device__ __forceinline__ void candidates(int y, float r, int out, int& lo, int& hi) {
    if (r <= 0.f) {  // out == 1: the single output samples source 0
        lo = 0;
        hi = (y == 0) ? 0 : -1;
        return;
    }
    lo = max(0, (int)floorf((float)(y - 1) / r) - 1);
    hi = min(out - 1, (int)ceilf((float)(y + 1) / r) + 1);
}

// one thread per element of the [C, H, W] gradient, x fastest
__global__ void __launch_bounds__(256) resize_bwd_kernel(ResizeGeom g, const float* __restrict__ dout,
                                                         float* __restrict__ dfm) {
    const size_t n = (size_t)g.C * g.H * g.W;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = (int)(i % g.W);
    const int y = (int)((i / g.W) % g.H);
    const int c = (int)(i / ((size_t)g.W * g.H));
    int oy_lo, oy_hi, ox_lo, ox_hi;
    candidates(y, g.ry, g.Hg, oy_lo, oy_hi);
    candidates(x, g.rx, g.Wg, ox_lo, ox_hi);
    const float* p = dout + (size_t)c * g.Hg * g.Wg;
    float acc = 0.f;
    for (int oy = oy_lo; oy <= oy_hi; oy++) {
        int y0, y1;
        float ly0, ly1;
        src_of(oy, g.ry, g.H, y0, y1, ly0, ly1);
        const float wy = (y0 == y ? ly0 : 0.f) + (y1 == y ? ly1 : 0.f);
        if (wy == 0.f) continue;
        float row = 0.f;
        for (int ox = ox_lo; ox <= ox_hi; ox++) {
            int x0, x1;
            float lx0, lx1;
            src_of(ox, g.rx, g.W, x0, x1, lx0, lx1);
            const float wx = (x0 == x ? lx0 : 0.f) + (x1 == x ? lx1 : 0.f);
            if (wx != 0.f) row += wx * __ldg(p + (size_t)oy * g.Wg + ox);
        }
        acc += wy * row;
    }
    dfm[i] = acc;
}

ResizeGeom make_geom(int C, int H, int W, int Hg, int Wg) {
    ResizeGeom g;
    g.C = C; g.H = H; g.W = W; g.Hg = Hg; g.Wg = Wg;
    g.ry = Hg > 1 ? (float)(H - 1) / (float)(Hg - 1) : 0.f;
    g.rx = Wg > 1 ? (float)(W - 1) / (float)(Wg - 1) : 0.f;
    return g;
}

}  // namespace

cudaError_t launch_feature_resize_fwd(int C, int H, int W, int Hg, int Wg, const float* fm, const float* gt,
                                      float grad_scale, float* out, float* loss_sum, cudaStream_t s) {
    const size_t n = (size_t)C * Hg * Wg;
    if (n == 0) return cudaSuccess;
    resize_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(make_geom(C, H, W, Hg, Wg), fm, gt, grad_scale, out, loss_sum);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_feature_resize_bwd(int C, int H, int W, int Hg, int Wg, const float* dout, float* dfm, cudaStream_t s) {
    const size_t n = (size_t)C * H * W;
    if (n == 0) return cudaSuccess;
    resize_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(make_geom(C, H, W, Hg, Wg), dout, dfm);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace f3dgs
