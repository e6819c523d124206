// Two-pass composite (experimental, opt-in with F3DGS_SPLIT=1; the fused kernels of composite_fwd.cu / composite_bwd.cu
// stay the default until this path has been validated and timed on the GPU).
//
// The fused kernels couple three warp roles through shared-memory rings; they run at ~50 % of their instruction-issue
// floor because the roles wait for each other, while HBM sits at ~10 % utilisation.  This path spends some of that idle
// bandwidth to decouple them:
//
//   alpha pass    composite_fwd_kernel<0, 1, EMIT> (composite_fwd.cu): the C = 0 forward kernel -- colour, depth, final_T,
//                 n_contrib exactly as before -- which also appends, per (tile, 8x4 block), one list entry per instance
//                 that blended at least one pixel of the block: {Gaussian id, pixel mask} + the 32 blend weights
//                 w = alpha * T (136 bytes per entry, ~6.0 M entries = 0.8 GB per view at config 3, front to back).
//                 Lists need no counting pass: block b of tile t owns entries [8*range.x + b*len, ... + len),
//                 len = range.y - range.x (an instance of the tile list appears at most once per block).
//   feature pass  feature_fwd_kernel<CH> (here): every warp is an independent worker that pulls (tile, channel chunk, block)
//                 items from an atomic counter, streams the block's list through a double-buffered cp.async ring
//                 (weights + the instances' feature rows, 8 entries per step) and accumulates the block's
//                 32 pixels x 4 channels per lane with the same FFMA2 quad loop and in the same order as the fused kernel,
//                 so the feature map is bit-identical.  No inter-warp synchronisation at all; 12 warps per SM.
//   backward      the geometric gradients come from the C = 0 backward kernel (composite_bwd.cu, unchanged); the
//                 feature gradient  dL/df[g] += sum_pixels w * dL/dO  needs only the forward's lists:
//                 feature_bwd_kernel<CH> streams them once more (weights only) and issues one red.global.add.v4 per lane.
//                 It uses the forward's w = alpha*T instead of the backward's unwound T (a few ulp closer to exact).
//
// Channel counts above 128 reuse the same lists for every 128-channel chunk: the alpha evaluation is no longer repeated
// per chunk as in the fused kernels.
// Reference semantics: forward.cu:261-396 (feature accumulation :362-368), backward.cu:565-575 (feature gradient).
#include "composite_common.cuh"

namespace f3dgs {

constexpr int kListChunk = 8;   // list entries staged per pipeline step
constexpr int kFeatWarps = 4;   // independent worker warps per CTA

template <int CH, bool WITH_ROWS>
struct alignas(128) FeatSmem {  // per warp
    float w[2][kListChunk][32];
    float f[WITH_ROWS ? 2 : 1][WITH_ROWS ? kListChunk : 1][WITH_ROWS ? CH : 4];
};

struct FeatArgs {
    const uint2* ranges;
    const float* list_w;
    const uint2* list_meta;
    const uint32_t* list_cnt;
    const float* features;      // forward: [P, C]
    float* out_feature;         // forward: [C, H, W]
    const float* dL_dfeat_pix;  // backward: [C, H, W]
    float* dL_dfeature;         // backward: [P, C]
    int* work_counter;
    int W, H, C, tiles_x, num_tiles, chunks;
    int vec;  // bit0: feature / gradient rows are 16-byte aligned and C % 4 == 0; bit1: 128-bit image rows; bit2: 256-bit
};

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Decode a work item.  Blocks of one tile are neighbours in the item order, so the workers that run at the same time
// mostly share their instances' feature rows in L2.
struct ItemPos {
    int tile, chunk, b, bx0, by0;
};
__device__ __forceinline__ ItemPos decode_item(int item, const FeatArgs& a) {
    ItemPos p;
    p.b = item & (kBlocksPerTile - 1);
    const int tc = item / kBlocksPerTile;
    p.chunk = tc % a.chunks;
    p.tile = tc / a.chunks;
    const int tile_x = p.tile % a.tiles_x, tile_y = p.tile / a.tiles_x;
    p.bx0 = tile_x * 16 + (p.b & 1) * 8;
    p.by0 = tile_y * 16 + (p.b >> 1) * 4;
    return p;
}

// ------------------------------------------------------------------------------------------------ forward
template <int CH>
__global__ void __launch_bounds__(kFeatWarps * 32, 3) feature_fwd_kernel(const FeatArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    FeatSmem<CH, true>& sm = reinterpret_cast<FeatSmem<CH, true>*>(smem_raw)[warp];
    constexpr int LPR = CH / 4;   // lanes per feature row
    constexpr int G = 32 / LPR;   // lane groups sharing the 32 pixels
    constexpr int NQ = 8 / G;     // 2x2 quads per lane
    const int grp = lane / LPR, cl = lane % LPR;
    const int W = a.W, H = a.H, C = a.C;
    const size_t HW = (size_t)H * W;

    {  // staging rows start defined (rows shorter than CH leave their tail untouched)
        float4* p = reinterpret_cast<float4*>(&sm.f[0][0][0]);
        for (int i = lane; i < 2 * kListChunk * CH / 4; i += 32) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
    }

    const int items = a.num_tiles * a.chunks * kBlocksPerTile;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(a.work_counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= items) break;
        const ItemPos ip = decode_item(item, a);
        // loaded from uniform addresses, but only a shuffle tells ptxas that the values are warp-uniform (uniform loop
        // trip counts and branches: no reconvergence pairs around the quad tests)
        const uint32_t rx = __shfl_sync(0xffffffffu, a.ranges[ip.tile].x, 0);
        const uint32_t ry = __shfl_sync(0xffffffffu, a.ranges[ip.tile].y, 0);
        const size_t base = 8 * (size_t)rx + (size_t)ip.b * (ry - rx);
        const uint32_t n = __shfl_sync(0xffffffffu, a.list_cnt[(size_t)ip.tile * kBlocksPerTile + ip.b], 0);
        const int chunk_off = ip.chunk * CH;
        const int row_floats = min(CH, C - chunk_off);
        const uint32_t nch = (n + kListChunk - 1) / kListChunk;

        float2 acc2[NQ][2][4];  // [quad][pixel pair (row of the 2x2 quad)][channel], as in composite_fwd.cu
#pragma unroll
        for (int q = 0; q < NQ; q++)
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc2[q][r][c] = make_float2(0.f, 0.f);

        auto load_meta = [&](uint32_t c) -> uint2 {
            const uint32_t e = c * kListChunk + lane;
            return (lane < kListChunk && e < n) ? __ldg(&a.list_meta[base + e]) : make_uint2(0u, 0u);
        };
        auto issue = [&](uint32_t c, int buf, uint2 m) {  // stage chunk c: weight rows + the instances' feature rows
            const uint32_t cnt = min((uint32_t)kListChunk, n - c * kListChunk);
            const float* wsrc = a.list_w + (base + (size_t)c * kListChunk) * 32;
            for (uint32_t j = lane; j < cnt * 8; j += 32) cp_async16(&sm.w[buf][0][0] + j * 4, wsrc + j * 4);
            for (uint32_t i = 0; i < cnt; i++) {
                const uint32_t gid = __shfl_sync(0xffffffffu, m.x, i);
                const float* src = a.features + (size_t)gid * C + chunk_off;
                if (a.vec & 1) {
                    if (lane * 4 < row_floats) cp_async16(&sm.f[buf][i][lane * 4], src + lane * 4);
                } else {
                    for (int c2 = lane; c2 < row_floats; c2 += 32) sm.f[buf][i][c2] = __ldg(src + c2);
                }
            }
            cp_async_commit();
        };

        uint2 m_cur = load_meta(0), m_nxt = load_meta(1);
        if (nch > 0) issue(0, 0, m_cur);
        for (uint32_t c = 0; c < nch; c++) {
            const int buf = c & 1;
            const uint2 m_nn = load_meta(c + 2);
            if (c + 1 < nch) {
                issue(c + 1, buf ^ 1, m_nxt);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const uint32_t cnt = min((uint32_t)kListChunk, n - c * kListChunk);
            for (uint32_t i = 0; i < cnt; i++) {
                const uint32_t pm = __shfl_sync(0xffffffffu, m_cur.y, i);
                const float4 f = *reinterpret_cast<const float4*>(&sm.f[buf][i][cl * 4]);
                const float fc[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) {
                    const int q = qi * G + grp;
                    if ((pm >> (4 * q)) & 0xFu) {
                        const float4 w4 = *reinterpret_cast<const float4*>(&sm.w[buf][i][4 * q]);
                        const float2 w01 = make_float2(w4.x, w4.y), w23 = make_float2(w4.z, w4.w);
#pragma unroll
                        for (int ch = 0; ch < 4; ch++) {
                            const float2 fb = make_float2(fc[ch], fc[ch]);
                            acc2[qi][0][ch] = __ffma2_rn(fb, w01, acc2[qi][0][ch]);
                            acc2[qi][1][ch] = __ffma2_rn(fb, w23, acc2[qi][1][ch]);
                        }
                    }
                }
            }
            __syncwarp();  // every lane is done with `buf` before the next step refills it
            m_cur = m_nxt;
            m_nxt = m_nn;
        }

        // ---- write the block's 32 pixels x CH channels (pixel i of quad q: x = (q&3)*2 + (i&1), y = (q>>2)*2 + (i>>1))
#define ACCF(q, i, c) (((i) & 1) ? acc2[q][(i) >> 1][c].y : acc2[q][(i) >> 1][c].x)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int ch = chunk_off + cl * 4 + c;
            if (ch < C) {  // structured ifs only (no `continue`): the warp provably reconverges before the next item
                float* plane = a.out_feature + (size_t)ch * HW;
                if (G == 1 && (a.vec & 4) && ip.bx0 + 8 <= W) {
#pragma unroll
                    for (int y = 0; y < 4; y++) {
                        const int yy = ip.by0 + y;
                        if (yy < H) {
                            const int qa = (y >> 1) * 4, i0 = (y & 1) * 2;
                            st_na_f8(plane + (size_t)yy * W + ip.bx0,
                                     make_float4(ACCF(qa % NQ, i0, c), ACCF(qa % NQ, i0 + 1, c), ACCF((qa + 1) % NQ, i0, c),
                                                 ACCF((qa + 1) % NQ, i0 + 1, c)),
                                     make_float4(ACCF((qa + 2) % NQ, i0, c), ACCF((qa + 2) % NQ, i0 + 1, c),
                                                 ACCF((qa + 3) % NQ, i0, c), ACCF((qa + 3) % NQ, i0 + 1, c)));
                        }
                    }
                } else {
#pragma unroll
                    for (int qi = 0; qi < NQ; qi++) {
                        const int q = qi * G + grp;
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int xx = ip.bx0 + (q & 3) * 2 + (i & 1), yy = ip.by0 + (q >> 2) * 2 + (i >> 1);
                            if (xx < W && yy < H) plane[(size_t)yy * W + xx] = ACCF(qi, i, c);
                        }
                    }
                }
            }
        }
        __syncwarp();
#undef ACCF
    }
}

// ------------------------------------------------------------------------------------------------ backward
template <int CH>
__global__ void __launch_bounds__(kFeatWarps * 32, 3) feature_bwd_kernel(const FeatArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    FeatSmem<CH, false>& sm = reinterpret_cast<FeatSmem<CH, false>*>(smem_raw)[warp];
    constexpr int LPR = CH / 4;
    constexpr int G = 32 / LPR;
    constexpr int NQ = 8 / G;
    const int grp = lane / LPR, cl = lane % LPR;
    const int W = a.W, H = a.H, C = a.C;
    const size_t HW = (size_t)H * W;

    const int items = a.num_tiles * a.chunks * kBlocksPerTile;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(a.work_counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= items) break;
        const ItemPos ip = decode_item(item, a);
        // loaded from uniform addresses, but only a shuffle tells ptxas that the values are warp-uniform (uniform loop
        // trip counts and branches: no reconvergence pairs around the quad tests)
        const uint32_t rx = __shfl_sync(0xffffffffu, a.ranges[ip.tile].x, 0);
        const uint32_t ry = __shfl_sync(0xffffffffu, a.ranges[ip.tile].y, 0);
        const size_t base = 8 * (size_t)rx + (size_t)ip.b * (ry - rx);
        const uint32_t n = __shfl_sync(0xffffffffu, a.list_cnt[(size_t)ip.tile * kBlocksPerTile + ip.b], 0);
        if (n == 0) continue;
        const int ch0 = ip.chunk * CH + cl * 4;
        const uint32_t nch = (n + kListChunk - 1) / kListChunk;

        auto load_meta = [&](uint32_t c) -> uint2 {
            const uint32_t e = c * kListChunk + lane;
            return (lane < kListChunk && e < n) ? __ldg(&a.list_meta[base + e]) : make_uint2(0u, 0u);
        };
        auto issue = [&](uint32_t c, int buf) {
            const uint32_t cnt = min((uint32_t)kListChunk, n - c * kListChunk);
            const float* wsrc = a.list_w + (base + (size_t)c * kListChunk) * 32;
            for (uint32_t j = lane; j < cnt * 8; j += 32) cp_async16(&sm.w[buf][0][0] + j * 4, wsrc + j * 4);
            cp_async_commit();
        };
        uint2 m_cur = load_meta(0);
        issue(0, 0);

        // upstream gradient of the block's 32 pixels x 4 channels: [quad][pixel pair][channel], pairs as in composite_bwd.cu
        float2 dO2[NQ][2][4];
#define DOB(q, i, c) (((i) & 1) ? dO2[q][(i) >> 1][c].y : dO2[q][(i) >> 1][c].x)
#pragma unroll
        for (int q = 0; q < NQ; q++)
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) dO2[q][r][c] = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int ch = ch0 + c;
            if (ch >= C) continue;
            const float* plane = a.dL_dfeat_pix + (size_t)ch * HW;
            if (G == 1 && (a.vec & 2)) {
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    const int yy = ip.by0 + y;
                    if (yy >= H) continue;
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        const int xx = ip.bx0 + half * 4;
                        if (xx >= W) continue;
                        const int qa = (y >> 1) * 4 + half * 2, i0 = (y & 1) * 2;
                        const float4 v = ld_nc_f4(plane + (size_t)yy * W + xx);
                        DOB(qa % NQ, i0, c) = v.x;
                        DOB(qa % NQ, i0 + 1, c) = v.y;
                        DOB((qa + 1) % NQ, i0, c) = v.z;
                        DOB((qa + 1) % NQ, i0 + 1, c) = v.w;
                    }
                }
            } else {
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) {
                    const int q = qi * G + grp;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int xx = ip.bx0 + (q & 3) * 2 + (i & 1), yy = ip.by0 + (q >> 2) * 2 + (i >> 1);
                        if (xx < W && yy < H) DOB(qi, i, c) = __ldg(plane + (size_t)yy * W + xx);
                    }
                }
            }
        }
#undef DOB

        for (uint32_t c = 0; c < nch; c++) {
            const int buf = c & 1;
            const uint2 m_nxt = load_meta(c + 1);
            if (c + 1 < nch) {
                issue(c + 1, buf ^ 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const uint32_t cnt = min((uint32_t)kListChunk, n - c * kListChunk);
            for (uint32_t i = 0; i < cnt; i++) {
                const uint32_t gid = __shfl_sync(0xffffffffu, m_cur.x, i);
                const uint32_t pm = __shfl_sync(0xffffffffu, m_cur.y, i);
                float2 gp[4];  // per channel: (sum over even pixel columns, sum over odd pixel columns)
#pragma unroll
                for (int ch = 0; ch < 4; ch++) gp[ch] = make_float2(0.f, 0.f);
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) {
                    const int q = qi * G + grp;
                    if ((pm >> (4 * q)) & 0xFu) {
                        const float4 w4 = *reinterpret_cast<const float4*>(&sm.w[buf][i][4 * q]);
                        const float2 w01 = make_float2(w4.x, w4.y), w23 = make_float2(w4.z, w4.w);
#pragma unroll
                        for (int ch = 0; ch < 4; ch++) gp[ch] = __ffma2_rn(w01, dO2[qi][0][ch], gp[ch]);
#pragma unroll
                        for (int ch = 0; ch < 4; ch++) gp[ch] = __ffma2_rn(w23, dO2[qi][1][ch], gp[ch]);
                    }
                }
                float g0 = gp[0].x + gp[0].y, g1 = gp[1].x + gp[1].y, g2 = gp[2].x + gp[2].y, g3 = gp[3].x + gp[3].y;
#pragma unroll
                for (int o = LPR; o < 32; o <<= 1) {
                    g0 += __shfl_xor_sync(0xffffffffu, g0, o);
                    g1 += __shfl_xor_sync(0xffffffffu, g1, o);
                    g2 += __shfl_xor_sync(0xffffffffu, g2, o);
                    g3 += __shfl_xor_sync(0xffffffffu, g3, o);
                }
                if (grp == 0 && ch0 < C) {
                    float* dst = a.dL_dfeature + (size_t)gid * C + ch0;
                    if (a.vec & 1) {
                        red_add_f4(dst, make_float4(g0, g1, g2, g3));
                    } else {
                        red_add_f1(dst, g0);
                        if (ch0 + 1 < C) red_add_f1(dst + 1, g1);
                        if (ch0 + 2 < C) red_add_f1(dst + 2, g2);
                        if (ch0 + 3 < C) red_add_f1(dst + 3, g3);
                    }
                }
            }
            __syncwarp();
            m_cur = m_nxt;
        }
    }
}

// ------------------------------------------------------------------------------------------------ launchers
static int workers_grid() {
    static int sms_of_device[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148 * 3;
    if (sms_of_device[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms_of_device[dev] = n > 0 ? n : 148;
    }
    return sms_of_device[dev] * 3;  // __launch_bounds__(128, 3): three CTAs of four workers per SM
}

template <int CH>
static cudaError_t launch_feat_fwd_t(const FeatArgs& a, cudaStream_t s) {
    const size_t smem = kFeatWarps * sizeof(FeatSmem<CH, true>);
    static_assert(kFeatWarps * sizeof(FeatSmem<CH, true>) <= 48 * 1024, "feature pass staging must fit the default limit");
    const int items = a.num_tiles * a.chunks * kBlocksPerTile;
    const int grid = min((items + kFeatWarps - 1) / kFeatWarps, workers_grid());
    feature_fwd_kernel<CH><<<grid, kFeatWarps * 32, smem, s>>>(a);
    g_launches++;
    return cudaGetLastError();
}
template <int CH>
static cudaError_t launch_feat_bwd_t(const FeatArgs& a, cudaStream_t s) {
    const size_t smem = kFeatWarps * sizeof(FeatSmem<CH, false>);
    const int items = a.num_tiles * a.chunks * kBlocksPerTile;
    const int grid = min((items + kFeatWarps - 1) / kFeatWarps, workers_grid());
    feature_bwd_kernel<CH><<<grid, kFeatWarps * 32, smem, s>>>(a);
    g_launches++;
    return cudaGetLastError();
}

static int feat_ch(int C) { return C <= 32 ? 32 : (C <= 64 ? 64 : 128); }

cudaError_t launch_feature_fwd(const ViewParams& vp, const uint2* ranges, const float* list_w, const uint2* list_meta,
                               const uint32_t* list_cnt, const float* features, float* out_feature, int* work_counter,
                               cudaStream_t s) {
    FeatArgs a;
    a.ranges = ranges; a.list_w = list_w; a.list_meta = list_meta; a.list_cnt = list_cnt;
    a.features = features; a.out_feature = out_feature; a.dL_dfeat_pix = nullptr; a.dL_dfeature = nullptr;
    a.work_counter = work_counter;
    a.W = vp.W; a.H = vp.H; a.C = vp.C; a.tiles_x = (int)vp.grid_x; a.num_tiles = (int)(vp.grid_x * vp.grid_y);
    const int CH = feat_ch(vp.C);
    a.chunks = (vp.C + CH - 1) / CH;
    a.vec = 0;
    if (vp.C % 4 == 0 && (reinterpret_cast<uintptr_t>(features) & 15) == 0) a.vec |= 1;
    if (vp.W % 4 == 0 && (reinterpret_cast<uintptr_t>(out_feature) & 15) == 0) a.vec |= 2;
    if (vp.W % 8 == 0 && (reinterpret_cast<uintptr_t>(out_feature) & 31) == 0) a.vec |= 4;
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    if (CH == 32) return launch_feat_fwd_t<32>(a, s);
    if (CH == 64) return launch_feat_fwd_t<64>(a, s);
    return launch_feat_fwd_t<128>(a, s);
}

cudaError_t launch_feature_bwd(const ViewParams& vp, const uint2* ranges, const float* list_w, const uint2* list_meta,
                               const uint32_t* list_cnt, const float* dL_dfeat_pix, float* dL_dfeature,
                               int* work_counter, cudaStream_t s) {
    FeatArgs a;
    a.ranges = ranges; a.list_w = list_w; a.list_meta = list_meta; a.list_cnt = list_cnt;
    a.features = nullptr; a.out_feature = nullptr; a.dL_dfeat_pix = dL_dfeat_pix; a.dL_dfeature = dL_dfeature;
    a.work_counter = work_counter;
    a.W = vp.W; a.H = vp.H; a.C = vp.C; a.tiles_x = (int)vp.grid_x; a.num_tiles = (int)(vp.grid_x * vp.grid_y);
    const int CH = feat_ch(vp.C);
    a.chunks = (vp.C + CH - 1) / CH;
    a.vec = 0;
    if (vp.C % 4 == 0 && (reinterpret_cast<uintptr_t>(dL_dfeature) & 15) == 0) a.vec |= 1;
    if (vp.W % 4 == 0 && (reinterpret_cast<uintptr_t>(dL_dfeat_pix) & 15) == 0) a.vec |= 2;
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    if (CH == 32) return launch_feat_bwd_t<32>(a, s);
    if (CH == 64) return launch_feat_bwd_t<64>(a, s);
    return launch_feat_bwd_t<128>(a, s);
}

}  // namespace f3dgs
