// Forward tile composite: front-to-back alpha blend of RGB, depth and a C-wide feature vector.
// Reference: forward.cu:261-396 (renderCUDA<3>), semantics restated in SURVEY.md A.4.
//
// Per pixel the arithmetic that decides WHAT is blended (power, alpha, T, the 1/255 and 1e-4
// tests, n_contrib, final_T) and the RGB/depth accumulation are the reference's expressions
// compiled by the same compiler (see common.cuh), so those outputs are bit-identical.  The feature
// accumulation uses acc = fma(f, alpha*T, acc) instead of fma(T, alpha*f, acc): one rounding
// differs per term (<= 1 ulp of the term), which halves the FMA-pipe work of the hot loop.
//
// Roles (composite_common.cuh): producer -> alpha warps -> feature warps, persistent over tiles.
//   alpha warp b   lane = pixel of block b.  For every staged instance whose footprint reaches
//                  the block: alpha, T update, RGB/depth accumulate; publishes w = alpha*T as a
//                  [instance][pixel] tile + a 32-bit "which pixels blended" mask per instance.
//   feature warp b lane = float4 of channels.  acc[pixel][4] += w[pixel] * f[4] for the pixels in
//                  the mask, 2x2-pixel quads at a time: one broadcast LDS.128 of weights feeds
//                  16 FFMAs; the feature float4 is loaded once per instance.  All 32 pixels x 4
//                  channels (= 128 accumulators per lane at CH = 128) live in registers.
// Channel counts above 128 are split into chunks of 128 (extra work items); chunk 0 also writes
// colour, depth, final_T and n_contrib.
#include <cstdlib>

#include "composite_common.cuh"

namespace f3dgs {

struct FwdArgs {
    ProducerArgs pa;
    const float* bg;
    float* final_T;
    uint32_t* n_contrib;
    float* out_color;
    float* out_feature;
    float* out_depth;
    int vec_store;
};

template <int CH, int BPA>
__global__ void __launch_bounds__(Layout<BPA>::kThreads, 1) composite_fwd_kernel(const FwdArgs args) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RingV2<CH>& ring = *reinterpret_cast<RingV2<CH>*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = args.pa.W, H = args.pa.H, C = args.pa.C;
    const size_t HW = (size_t)H * W;

    using L = Layout<BPA>;
    ring_init(ring, CH > 0 ? L::kAlphaWarps + kBlocksPerTile : L::kAlphaWarps, CH > 0);
    __syncthreads();

    // ======================================================================== producer group
    if (warp < kAlphaWarp0) {
        reg_dec<L::kRegsProducer>();
        if (warp == kProducerWarp) producer_loop<CH, false>(ring, args.pa);
        return;
    }

    // ======================================================================== alpha warps
    if (warp < L::kFeatWarp0) {
        reg_dec<L::kRegsAlpha>();
        const int a = warp - kAlphaWarp0;  // owns blocks BPA*a .. BPA*a + BPA-1
        int s = 0, j = 0;
        uint32_t parity = 0, wparity = 1;  // wempty: fresh barrier falls through on parity 1
        float T[BPA], Cr[BPA], Cg[BPA], Cb[BPA], Dp[BPA], pxf[BPA], pyf[BPA], fbx0[BPA], fby0[BPA];
        uint32_t last_contrib[BPA];
        int px[BPA], py[BPA], chunk = 0;
        bool done[BPA], inside[BPA], blk_done[BPA];
#pragma unroll
        for (int bi = 0; bi < BPA; bi++) {
            T[bi] = 1.f; Cr[bi] = Cg[bi] = Cb[bi] = Dp[bi] = 0.f; pxf[bi] = pyf[bi] = fbx0[bi] = fby0[bi] = 0.f;
            last_contrib[bi] = 0; px[bi] = py[bi] = 0; done[bi] = true; inside[bi] = false; blk_done[bi] = true;
        }
        for (;;) {
            mbar_wait(&ring.full[s], parity);
            Stage<CH>& st = ring.stage[s];
            const uint32_t n = st.n, last = st.last, first = st.first;
            const int work = st.work;
            if (work < 0) break;
            if (first) {
                const int tile = work / args.pa.chunks;
                chunk = work - tile * args.pa.chunks;
                const int tile_x = tile % args.pa.tiles_x, tile_y = tile / args.pa.tiles_x;
#pragma unroll
                for (int bi = 0; bi < BPA; bi++) {
                    const int b = BPA * a + bi;
                    const int bx0 = tile_x * 16 + (b & 1) * 8, by0 = tile_y * 16 + (b >> 1) * 4;
                    px[bi] = bx0 + lane_px(lane);
                    py[bi] = by0 + lane_py(lane);
                    inside[bi] = px[bi] < W && py[bi] < H;
                    pxf[bi] = (float)px[bi]; pyf[bi] = (float)py[bi];
                    fbx0[bi] = (float)bx0; fby0[bi] = (float)by0;
                    T[bi] = 1.f; Cr[bi] = Cg[bi] = Cb[bi] = Dp[bi] = 0.f;
                    last_contrib[bi] = 0;
                    done[bi] = !inside[bi];
                    blk_done[bi] = __all_sync(0xffffffffu, done[bi]);
                    if (blk_done[bi] && lane == 0) atomicOr(&ring.done_mask[work % kDoneSlots], 1u << b);
                }
            }
#pragma unroll
            for (int bi = 0; bi < BPA; bi++) {
                const int b = BPA * a + bi;
                WSlot* ws = &ring.ws[b][j];
                if (CH > 0) mbar_wait(&ring.wempty[b][j], wparity);
                uint32_t km = 0;
                if (!blk_done[bi] && n > 0) {
                    bool hit = false;
                    if (lane < n) {
                        const float4 r0 = st.rec0[lane];
                        hit = (r0.x + r0.z >= fbx0[bi]) && (r0.x - r0.z <= fbx0[bi] + 7.f) &&
                              (r0.y + r0.w >= fby0[bi]) && (r0.y - r0.w <= fby0[bi] + 3.f);
                    }
                    uint32_t am = __ballot_sync(0xffffffffu, hit);
                    while (am) {
                        // Up to 4 instances per trip.  Everything is branch-free so that the four alpha
                        // evaluations (LDS -> power -> expf) interleave in the pipeline; only the short
                        // T / done recurrence is serial.
                        int kk[4];
                        bool vk[4];
                        float al[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            vk[u] = am != 0;
                            kk[u] = vk[u] ? (__ffs(am) - 1) : 0;
                            am &= am - 1;
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const float4 r0 = st.rec0[kk[u]];
                            const float4 r1 = st.rec1[kk[u]];
                            // same expression trees as reference forward.cu:340-351 (see common.cuh)
                            const float dx = r0.x - pxf[bi], dy = r0.y - pyf[bi];
                            const float power = -0.5f * (r1.x * dx * dx + r1.z * dy * dy) - r1.y * dx * dy;
                            const float av = fminf(0.99f, r1.w * expf(power));
                            al[u] = (vk[u] && !(power > 0.0f) && !(av < 1.0f / 255.0f)) ? av : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (!vk[u]) break;  // warp-uniform, only in the last trip
                            const float4 r2 = st.rec2[kk[u]];
                            const uint32_t lp = st.listpos[kk[u]];
                            const float alpha = al[u];
                            const float test_T = T[bi] * (1 - alpha);
                            const bool act = !done[bi] && alpha > 0.f;
                            const bool stop = act && (test_T < 0.0001f);  // reference: done = true, not blended
                            const bool blend = act && !stop;
                            done[bi] = done[bi] || stop;
                            const float wgt = blend ? alpha * T[bi] : 0.f;
                            const float nCr = Cr[bi] + r2.x * alpha * T[bi];  // reference forward.cu:362-368
                            const float nCg = Cg[bi] + r2.y * alpha * T[bi];
                            const float nCb = Cb[bi] + r2.z * alpha * T[bi];
                            const float nDp = Dp[bi] + r2.w * (alpha * T[bi]);
                            Cr[bi] = blend ? nCr : Cr[bi];
                            Cg[bi] = blend ? nCg : Cg[bi];
                            Cb[bi] = blend ? nCb : Cb[bi];
                            Dp[bi] = blend ? nDp : Dp[bi];
                            T[bi] = blend ? test_T : T[bi];
                            last_contrib[bi] = blend ? lp : last_contrib[bi];
                            const uint32_t pm = __ballot_sync(0xffffffffu, blend);
                            if (CH > 0) {
                                ws->w[kk[u]][lane] = wgt;
                                if (lane == 0) ws->pm[kk[u]] = pm;
                                km |= (pm ? 1u : 0u) << kk[u];
                            }
                        }
                    }
                    if (__all_sync(0xffffffffu, done[bi])) {
                        blk_done[bi] = true;
                        if (lane == 0) atomicOr(&ring.done_mask[work % kDoneSlots], 1u << b);
                    }
                }
                if (CH > 0) {
                    __syncwarp();
                    if (lane == 0) {
                        ws->km = km;
                        ws->last = last;
                        ws->first = first;
                        ws->work = work;
                        mbar_arrive(&ring.wfull[b][j]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.empty[s]);
            if (last && chunk == 0) {
#pragma unroll
                for (int bi = 0; bi < BPA; bi++)
                    if (inside[bi]) {
                        const size_t pix = (size_t)py[bi] * W + px[bi];
                        args.final_T[pix] = T[bi];
                        args.n_contrib[pix] = last_contrib[bi];
                        args.out_color[pix] = Cr[bi] + T[bi] * args.bg[0];  // reference forward.cu:389
                        args.out_color[HW + pix] = Cg[bi] + T[bi] * args.bg[1];
                        args.out_color[2 * HW + pix] = Cb[bi] + T[bi] * args.bg[2];
                        args.out_depth[pix] = Dp[bi];
                    }
            }
            if (++s == kStages) { s = 0; parity ^= 1; }
            if (CH > 0 && ++j == kWSlots) { j = 0; wparity ^= 1; }
        }
        if (CH > 0) {  // tell the feature warps of these blocks that the work is over
#pragma unroll
            for (int bi = 0; bi < BPA; bi++) {
                const int b = BPA * a + bi;
                mbar_wait(&ring.wempty[b][j], wparity);
                if (lane == 0) {
                    ring.ws[b][j].work = -1;
                    ring.ws[b][j].km = 0;
                    mbar_arrive(&ring.wfull[b][j]);
                }
            }
        }
        return;
    }

    // ======================================================================== feature warps
    if (CH == 0) return;
    reg_inc<L::kRegsFeature>();
    {
        constexpr int LPR = CH > 0 ? CH / 4 : 32;  // lanes per feature row
        constexpr int G = 32 / LPR;                // lane groups sharing the 32 pixels
        constexpr int NQ = 8 / G;                  // 2x2 quads per lane
        const int b = warp - L::kFeatWarp0;
        const int grp = lane / LPR, cl = lane % LPR;
        float acc[NQ][4][4];  // [quad][pixel in quad][channel]
#pragma unroll
        for (int q = 0; q < NQ; q++)
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[q][i][c] = 0.f;
        int s = 0, j = 0;
        uint32_t parity = 0, wparity = 0;
        for (;;) {
            mbar_wait(&ring.wfull[b][j], wparity);
            const WSlot& ws = ring.ws[b][j];
            const int work = ws.work;
            if (work < 0) break;
            uint32_t km = ws.km;
            const uint32_t last = ws.last;
            mbar_wait(&ring.full[s], parity);  // feature rows of this stage have landed
            const Stage<CH>& st = ring.stage[s];
            if (L::kPrefetchW) {
                if (km) {
                    // software pipeline over the blended instances: the next instance's mask and feature
                    // float4 are fetched while the current one is being accumulated; all weight quads of
                    // the current instance are loaded up front so no LDS sits between the FFMA blocks
                    int k = __ffs(km) - 1;
                    km &= km - 1;
                    uint32_t pm = ws.pm[k];
                    float4 f = *reinterpret_cast<const float4*>(&st.feat[k][cl * 4]);
                    for (;;) {
                        float4 w4[NQ];
#pragma unroll
                        for (int qi = 0; qi < NQ; qi++)
                            w4[qi] = *reinterpret_cast<const float4*>(&ws.w[k][4 * (qi * G + grp)]);
                        const bool more = km != 0;
                        const int kn = more ? (__ffs(km) - 1) : k;
                        km &= km - 1;
                        const uint32_t pmn = ws.pm[kn];
                        const float4 fn = *reinterpret_cast<const float4*>(&st.feat[kn][cl * 4]);
#pragma unroll
                        for (int qi = 0; qi < NQ; qi++) {
                            const int q = qi * G + grp;
                            if ((pm >> (4 * q)) & 0xFu) {
                                // a pixel that skipped this instance stored w = 0: adds exactly 0
                                const float wv[4] = {w4[qi].x, w4[qi].y, w4[qi].z, w4[qi].w};
#pragma unroll
                                for (int i = 0; i < 4; i++) {
                                    acc[qi][i][0] = fmaf(f.x, wv[i], acc[qi][i][0]);
                                    acc[qi][i][1] = fmaf(f.y, wv[i], acc[qi][i][1]);
                                    acc[qi][i][2] = fmaf(f.z, wv[i], acc[qi][i][2]);
                                    acc[qi][i][3] = fmaf(f.w, wv[i], acc[qi][i][3]);
                                }
                            }
                        }
                        if (!more) break;
                        k = kn; pm = pmn; f = fn;
                    }
                }
            } else {
                while (km) {
                    const int k = __ffs(km) - 1;
                    km &= km - 1;
                    const uint32_t pm = ws.pm[k];
                    const float4 f = *reinterpret_cast<const float4*>(&st.feat[k][cl * 4]);
#pragma unroll
                    for (int qi = 0; qi < NQ; qi++) {
                        const int q = qi * G + grp;
                        if ((pm >> (4 * q)) & 0xFu) {
                            const float4 w4 = *reinterpret_cast<const float4*>(&ws.w[k][4 * q]);
                            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                acc[qi][i][0] = fmaf(f.x, wv[i], acc[qi][i][0]);
                                acc[qi][i][1] = fmaf(f.y, wv[i], acc[qi][i][1]);
                                acc[qi][i][2] = fmaf(f.z, wv[i], acc[qi][i][2]);
                                acc[qi][i][3] = fmaf(f.w, wv[i], acc[qi][i][3]);
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&ring.wempty[b][j]);
                mbar_arrive(&ring.empty[s]);
            }
            if (last) {
                // ---- epilogue of this work item: write the block's 32 pixels x CH channels, reset
                const int tile = work / args.pa.chunks, chunk = work - tile * args.pa.chunks;
                const int tile_x = tile % args.pa.tiles_x, tile_y = tile / args.pa.tiles_x;
                const int bx0 = tile_x * 16 + (b & 1) * 8, by0 = tile_y * 16 + (b >> 1) * 4;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int ch = chunk * CH + cl * 4 + c;
                    if (ch < C) {
                        float* plane = args.out_feature + (size_t)ch * HW;
                        if (G == 1 && args.vec_store) {
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
                                    const float4 v =
                                        make_float4(acc[qa % NQ][i0][c], acc[qa % NQ][i0 + 1][c],
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
#pragma unroll
                for (int q = 0; q < NQ; q++)
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int c = 0; c < 4; c++) acc[q][i][c] = 0.f;
            }
            if (++s == kStages) { s = 0; parity ^= 1; }
            if (++j == kWSlots) { j = 0; wparity ^= 1; }
        }
    }
}

// Warp layout per kernel (see Layout<> in composite_common.cuh).  Measured at c3 on B200: the forward is
// fastest with 4 alpha warps x 2 blocks (feature warps get 184 registers and prefetch their weights), the
// backward with 8 alpha warps x 1 block (its alpha side carries the gradient reductions and dominates).
// F3DGS_BPA=1|2 overrides both for experiments.
int composite_layout_bpa(int default_bpa) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("F3DGS_BPA");
        forced = (e && (e[0] == '1' || e[0] == '2')) ? (e[0] - '0') : 0;
    }
    return forced ? forced : default_bpa;
}

template <int CH, int BPA>
static cudaError_t launch_fwd_t(const ViewParams& vp, const uint2* ranges, const uint32_t* point_list,
                                const SplatRec* rec, const float* features, const float* bg, float* final_T,
                                uint32_t* n_contrib, float* out_color, float* out_feature, float* out_depth,
                                int* work_counter, cudaStream_t s) {
    const size_t smem = sizeof(RingV2<CH>);
    static bool attr_set = false;
    static int num_sms = 0;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(composite_fwd_kernel<CH, BPA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        attr_set = true;
    }
    FwdArgs a;
    a.pa.ranges = ranges; a.pa.point_list = point_list; a.pa.rec = rec;
    a.pa.features = CH > 0 ? features : nullptr;
    a.pa.n_contrib = nullptr;
    a.pa.work_counter = work_counter;
    a.pa.W = vp.W; a.pa.H = vp.H; a.pa.C = vp.C;
    a.pa.tiles_x = (int)vp.grid_x;
    a.pa.num_tiles = (int)(vp.grid_x * vp.grid_y);
    a.pa.chunks = CH > 0 ? (vp.C + CH - 1) / CH : 1;
    a.pa.use_bulk = (CH > 0 && vp.C % 4 == 0 && (reinterpret_cast<uintptr_t>(features) & 15) == 0) ? 1 : 0;
    a.bg = bg; a.final_T = final_T; a.n_contrib = n_contrib;
    a.out_color = out_color; a.out_feature = out_feature; a.out_depth = out_depth;
    a.vec_store = (vp.W % 4 == 0 && (reinterpret_cast<uintptr_t>(out_feature) & 15) == 0) ? 1 : 0;
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    const int grid = min(a.pa.num_tiles * a.pa.chunks, num_sms > 0 ? num_sms : 148);
    composite_fwd_kernel<CH, BPA><<<grid, Layout<BPA>::kThreads, smem, s>>>(a);
    g_launches++;
    return cudaGetLastError();
}

#define F3DGS_FWD_DISPATCH(CHV)                                                                                   \
    (composite_layout_bpa(2) == 2                                                                                \
         ? launch_fwd_t<CHV, 2>(vp, ranges, point_list, rec, features, bg, final_T, n_contrib, out_color,       \
                                out_feature, out_depth, work_counter, s)                                        \
         : launch_fwd_t<CHV, 1>(vp, ranges, point_list, rec, features, bg, final_T, n_contrib, out_color,       \
                                out_feature, out_depth, work_counter, s))

cudaError_t launch_composite_fwd(const ViewParams& vp, const uint2* ranges, const uint32_t* point_list,
                                 const SplatRec* rec, const float* features, const float* bg,
                                 float* final_T, uint32_t* n_contrib, float* out_color,
                                 float* out_feature, float* out_depth, int* work_counter, cudaStream_t s) {
    if (vp.C == 0) return F3DGS_FWD_DISPATCH(0);
    if (vp.C <= 32) return F3DGS_FWD_DISPATCH(32);
    if (vp.C <= 64) return F3DGS_FWD_DISPATCH(64);
    return F3DGS_FWD_DISPATCH(128);
}

}  // namespace f3dgs
