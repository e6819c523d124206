// Forward tile composite: front-to-back alpha blend of RGB, depth and a C-wide feature vector.
// Reference: forward.cu:261-396 (renderCUDA<3>), semantics restated in SURVEY.md A.4.
//
// Per pixel the arithmetic that decides WHAT is blended (power, alpha, T, the 1/255 and 1e-4
// tests, n_contrib, final_T) and the RGB/depth accumulation follow the reference's operation
// sequence exactly (read off its PTX), so those outputs are bit-identical.  The feature
// accumulation uses acc = fma(f, alpha*T, acc) instead of fma(T, alpha*f, acc): one rounding
// differs per term (<= 1 ulp of the term), which halves the FMA-pipe work of the hot loop.
//
// Structure (see composite_common.cuh for the producer side):
//   alpha pass   lane = pixel.  For every staged instance whose footprint reaches the warp's
//                8x4 block: alpha, T update, RGB/depth accumulate, and the blend weight
//                w = alpha*T written to a per-warp [instance][pixel] tile in shared memory,
//                plus a 32-bit "which pixels blended" mask per instance (one ballot).
//   feature pass lane = float4 of channels.  acc[pixel][4] += w[pixel] * f[4] for the pixels in
//                the mask, 2x2-pixel quads at a time: one broadcast LDS.128 of weights feeds
//                16 FFMAs; the feature float4 is loaded once per instance.  All 32 pixels x 4
//                channels (= CH accumulators per lane for CH = 128) live in registers.
// Channel counts above 128 are split over gridDim.z chunks of 128; chunk 0 also writes colour,
// depth, final_T and n_contrib.
#include "composite_common.cuh"

namespace f3dgs {

constexpr int kFwdStages = 4;

template <int CH>
struct FwdSmem {
    Ring<CH, kFwdStages> ring;
    float w[kConsumerWarps][kStageEntries][32];  // blend weights [warp][instance][pixel]
};

template <int CH>
__global__ void __launch_bounds__(kBlockThreads, 1)
composite_fwd_kernel(int W, int H, int C, const uint2* __restrict__ ranges,
                     const uint32_t* __restrict__ point_list, const SplatRec* __restrict__ rec,
                     const float* __restrict__ features, const float* __restrict__ bg,
                     float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                     float* __restrict__ out_color, float* __restrict__ out_feature,
                     float* __restrict__ out_depth, int use_bulk, int vec_store) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FwdSmem<CH>& sm = *reinterpret_cast<FwdSmem<CH>*>(smem_raw);
    Ring<CH, kFwdStages>& ring = sm.ring;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y, chunk = blockIdx.z;
    const int chunk_off = chunk * CH;
    const uint2 range = ranges[tile_y * gridDim.x + tile_x];

    ring_init(ring);
    __syncthreads();

    if (warp == kConsumerWarps) {
        const int row_floats = CH > 0 ? min(CH, C - chunk_off) : 0;
        const float tx0 = (float)(tile_x * 16), ty0 = (float)(tile_y * 16);
        producer_loop<CH, kFwdStages, false>(ring, point_list, rec, features, C, chunk_off, row_floats,
                                             use_bulk != 0, range.x, range.y, range.y - range.x, tx0, ty0,
                                             tx0 + 15.f, ty0 + 15.f);
        return;
    }

    // ------------------------------------------------------------------ consumer warps
    constexpr int LPR = CH > 0 ? CH / 4 : 32;  // lanes per feature row
    constexpr int G = 32 / LPR;                // lane groups sharing the 32 pixels
    constexpr int NQ = 8 / G;                  // 2x2 quads per lane
    const int grp = lane / LPR, cl = lane % LPR;

    const int bx0 = tile_x * 16 + (warp & 1) * 8, by0 = tile_y * 16 + (warp >> 1) * 4;
    const int px = bx0 + lane_px(lane), py = by0 + lane_py(lane);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const float fbx0 = (float)bx0, fby0 = (float)by0, fbx1 = (float)(bx0 + 7), fby1 = (float)(by0 + 3);

    float T = 1.0f, Cr = 0.f, Cg = 0.f, Cb = 0.f, Dp = 0.f;
    uint32_t last_contrib = 0;
    bool done = !inside;
    bool warp_done = false;

    float acc[NQ][4][4];  // [quad][pixel in quad][channel]
#pragma unroll
    for (int q = 0; q < NQ; q++)
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[q][i][c] = 0.f;

    float(*wbuf)[32] = sm.w[warp];
    int s = 0;
    uint32_t parity = 0;
    if (__all_sync(0xffffffffu, done)) {
        warp_done = true;
        if (lane == 0) atomicOr(&ring.done_mask, 1u << warp);
    }

    while (true) {
        mbar_wait(&ring.full[s], parity);
        Stage<CH>& st = ring.stage[s];
        const uint32_t n = st.n;
        const uint32_t last = st.last;
        if (!warp_done && n > 0) {
            // which staged instances can touch this warp's 8x4 pixel block?
            bool hit = false;
            if (lane < n) {
                const float4 r0 = st.rec0[lane];
                hit = (r0.x + r0.z >= fbx0) && (r0.x - r0.z <= fbx1) && (r0.y + r0.w >= fby0) &&
                      (r0.y - r0.w <= fby1);
            }
            uint32_t am = __ballot_sync(0xffffffffu, hit);
            uint32_t mypm = 0;  // lane k: pixel mask of instance k
            while (am) {
                // up to 4 instances per trip: the alpha evaluations are independent (ILP),
                // only the T recurrence is serial
                int kk[4];
                float al[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    kk[u] = am ? (__ffs(am) - 1) : -1;
                    am &= am - 1;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    al[u] = 0.f;
                    if (kk[u] >= 0) {
                        const float4 r0 = st.rec0[kk[u]];
                        const float4 r1 = st.rec1[kk[u]];
                        // same expression trees as reference forward.cu:340-351 (see common.cuh)
                        const float dx = r0.x - pxf, dy = r0.y - pyf;
                        const float power = -0.5f * (r1.x * dx * dx + r1.z * dy * dy) - r1.y * dx * dy;
                        if (!(power > 0.0f)) {
                            const float a = fminf(0.99f, r1.w * expf(power));
                            if (!(a < 1.0f / 255.0f)) al[u] = a;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (kk[u] < 0) break;  // warp-uniform
                    float wgt = 0.f;
                    float alpha = al[u];
                    if (!done && alpha > 0.f) {
                        const float test_T = T * (1 - alpha);
                        if (test_T < 0.0001f) {
                            done = true;
                        } else {
                            const float4 r2 = st.rec2[kk[u]];
                            Cr += r2.x * alpha * T;  // reference forward.cu:362-368
                            Cg += r2.y * alpha * T;
                            Cb += r2.z * alpha * T;
                            wgt = alpha * T;
                            Dp += r2.w * wgt;
                            T = test_T;
                            last_contrib = st.listpos[kk[u]];
                        }
                    }
                    const uint32_t pm = __ballot_sync(0xffffffffu, wgt != 0.f);
                    if (CH > 0 && pm) {
                        wbuf[kk[u]][lane] = wgt;
                        if (lane == kk[u]) mypm = pm;
                    }
                }
            }
            if (CH > 0) {
                __syncwarp();
                uint32_t km = __ballot_sync(0xffffffffu, mypm != 0);
                while (km) {
                    const int k = __ffs(km) - 1;
                    km &= km - 1;
                    const uint32_t pm = __shfl_sync(0xffffffffu, mypm, k);
                    const float4 f = *reinterpret_cast<const float4*>(&st.feat[k][cl * 4]);
#pragma unroll
                    for (int qi = 0; qi < NQ; qi++) {
                        const int q = qi * G + grp;
                        if ((pm >> (4 * q)) & 0xFu) {
                            const float4 w4 = *reinterpret_cast<const float4*>(&wbuf[k][4 * q]);
                            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                // a pixel that skipped this instance stored w = 0: adds exactly 0
                                const float wi = wv[i];
                                acc[qi][i][0] = fmaf(f.x, wi, acc[qi][i][0]);
                                acc[qi][i][1] = fmaf(f.y, wi, acc[qi][i][1]);
                                acc[qi][i][2] = fmaf(f.z, wi, acc[qi][i][2]);
                                acc[qi][i][3] = fmaf(f.w, wi, acc[qi][i][3]);
                            }
                        }
                    }
                }
            }
            if (__all_sync(0xffffffffu, done)) {
                warp_done = true;
                if (lane == 0) atomicOr(&ring.done_mask, 1u << warp);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring.empty[s]);
        if (last) break;
        s++;
        if (s == kFwdStages) {
            s = 0;
            parity ^= 1;
        }
    }

    // ------------------------------------------------------------------ epilogue
    const size_t HW = (size_t)H * W;
    if (chunk == 0 && inside) {
        const size_t pix = (size_t)py * W + px;
        final_T[pix] = T;
        n_contrib[pix] = last_contrib;
        out_color[pix] = Cr + T * bg[0];  // reference forward.cu:389
        out_color[HW + pix] = Cg + T * bg[1];
        out_color[2 * HW + pix] = Cb + T * bg[2];
        out_depth[pix] = Dp;
    }
    if (CH > 0) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int ch = chunk_off + cl * 4 + c;
            if (ch >= C) continue;
            float* plane = out_feature + (size_t)ch * HW;
            if (G == 1 && vec_store) {
                // lane holds all 8 quads: rows of 8 pixels -> two 128-bit stores per row
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    const int yy = by0 + y;
                    if (yy >= H) continue;
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        const int xx = bx0 + half * 4;
                        if (xx >= W) continue;
                        const int qa = (y >> 1) * 4 + half * 2, i0 = (y & 1) * 2;
                        const float4 v = make_float4(acc[qa % NQ][i0][c], acc[qa % NQ][i0 + 1][c],
                                                     acc[(qa + 1) % NQ][i0][c], acc[(qa + 1) % NQ][i0 + 1][c]);
                        st_na_f4(plane + (size_t)yy * W + xx, v);
                    }
                }
            } else {
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) {
                    const int q = qi * G + grp;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int xx = bx0 + (q & 3) * 2 + (i & 1), yy = by0 + (q >> 2) * 2 + (i >> 1);
                        if (xx < W && yy < H) plane[(size_t)yy * W + xx] = acc[qi][i][c];
                    }
                }
            }
        }
    }
}

template <int CH>
static cudaError_t launch_fwd_t(const ViewParams& vp, const uint2* ranges, const uint32_t* point_list,
                                const SplatRec* rec, const float* features, const float* bg, float* final_T,
                                uint32_t* n_contrib, float* out_color, float* out_feature, float* out_depth,
                                cudaStream_t s) {
    const size_t smem = sizeof(FwdSmem<CH>);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(composite_fwd_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int chunks = CH > 0 ? (vp.C + CH - 1) / CH : 1;
    dim3 grid(vp.grid_x, vp.grid_y, chunks);
    const int use_bulk = (CH > 0 && vp.C % 4 == 0 && (reinterpret_cast<uintptr_t>(features) & 15) == 0) ? 1 : 0;
    const int vec_store = (vp.W % 4 == 0 && (reinterpret_cast<uintptr_t>(out_feature) & 15) == 0) ? 1 : 0;
    composite_fwd_kernel<CH><<<grid, kBlockThreads, smem, s>>>(vp.W, vp.H, vp.C, ranges, point_list, rec, features,
                                                              bg, final_T, n_contrib, out_color, out_feature,
                                                              out_depth, use_bulk, vec_store);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_composite_fwd(const ViewParams& vp, const uint2* ranges, const uint32_t* point_list,
                                 const SplatRec* rec, const float* features, const float* bg,
                                 float* final_T, uint32_t* n_contrib, float* out_color,
                                 float* out_feature, float* out_depth, cudaStream_t s) {
    if (vp.C == 0)
        return launch_fwd_t<0>(vp, ranges, point_list, rec, features, bg, final_T, n_contrib, out_color,
                               out_feature, out_depth, s);
    if (vp.C <= 32)
        return launch_fwd_t<32>(vp, ranges, point_list, rec, features, bg, final_T, n_contrib, out_color,
                                out_feature, out_depth, s);
    if (vp.C <= 64)
        return launch_fwd_t<64>(vp, ranges, point_list, rec, features, bg, final_T, n_contrib, out_color,
                                out_feature, out_depth, s);
    return launch_fwd_t<128>(vp, ranges, point_list, rec, features, bg, final_T, n_contrib, out_color,
                             out_feature, out_depth, s);
}

}  // namespace f3dgs
