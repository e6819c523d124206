// Forward tile composite with the feature contraction on the tensor cores (tcgen05, kind::tf32, 3xTF32 compensated).
// Reference: forward.cu:261-396 (renderCUDA<3>); semantics in SURVEY.md A.4.  Same outputs as composite_fwd.cu: colour,
// depth, final_T and n_contrib come from the same fp32 expressions (bit-identical to the reference build); the feature
// map differs from the fp32-pipe kernel by the rounding of the compensated TF32 product (tc_common.cuh, < 1e-6 relative).
//
// Per tile the feature map is  D[ch, px] = sum_k f[k, ch] * w[k, px]  over the tile's (culled) instance list, with
// w[k, px] = alpha*T where the pair blended and 0 elsewhere.  A persistent CTA (one per SM) runs five roles:
//
//   producer (1 warp)     producer_loop<> of composite_common.cuh: walks the tile's slice of the sorted list, drops instances
//                         whose footprint cannot reach the tile, fills the record ring (32 instances per stage).
//   alpha warps (8)       warp b = 8x4 pixel block b, lane = pixel: alpha, T recurrence, RGB / depth, final_T, n_contrib as
//                         in composite_fwd.cu; the blend weights go, split into hi / lo TF32 parts, into the B operand
//                         W[px block][k][32 px] (MN-major, SWIZZLE_128B_BASE32B) of the current operand stage (16 instances);
//                         rows of instances that miss the block stay zero.
//   convert warps (4)     fetch the stage's feature rows straight from HBM (one 512-byte row per warp load, 16 in
//                         flight), split them into hi / lo and write the A operand F[ch block][k][32 ch] (MN-major).
//   MMA issuer (1 lane)   per 8 instances three tcgen05.mma M=128 (channels) x N=256 (pixels) x K=8:
//                         F_hi*W_hi + F_hi*W_lo + F_lo*W_hi, fp32 accumulators in tensor memory (two 256-column buffers:
//                         the epilogue of tile t overlaps the MMAs of tile t+1).  tcgen05.commit releases operand stages.
//   epilogue warps (4)    tcgen05.ld: lane = channel, 32 columns = the 32 pixels of one block -> four 256-bit stores per
//                         block into the channel's CHW plane (full 32-byte sectors).
// Channel counts above 128 run as extra work items (chunks of 128 channels), as in composite_fwd.cu.
#include "composite_common.cuh"
#include "tc_common.cuh"

namespace f3dgs {

namespace {

constexpr int kKS = 16;         // instances per operand stage (two K=8 MMA steps)
constexpr int kOpStages = 4;    // operand ring depth
constexpr int kMmaWarp = 1;
constexpr int kAlpha0 = 4, kAlphaN = 8;
constexpr int kConv0 = 12, kConvN = 4;
constexpr int kEpi0 = 16, kEpiN = 4;   // kEpi0 % 4 == 0: epilogue warp e reads TMEM lanes 32e .. 32e+31
constexpr int kThreadsTc = (kEpi0 + kEpiN) * 32;

struct alignas(1024) OpStage {
    float Fhi[4][kKS][32];   // A operand: [channel block][instance][32 channels]
    float Flo[4][kKS][32];
    float Whi[8][kKS][32];   // B operand: [pixel block][instance][32 pixels]
    float Wlo[8][kKS][32];
};
static_assert(sizeof(OpStage) == 48 * 1024, "operand stage is 48 KB");

struct alignas(1024) TcSmem {
    OpStage op[kOpStages];
    RingSlim ring;                       // record ring (composite_common.cuh)
    uint64_t op_full[kOpStages];         // 8 alpha warps + 1 convert warp have written the stage
    uint64_t op_empty[kOpStages];        // the stage's MMAs have completed (tcgen05.commit)
    uint64_t tmem_full[2];               // a tile's accumulators are complete
    uint64_t tmem_empty[2];              // the epilogue has read them
    int32_t tile_work[2];
    uint32_t tmem_base;
};

struct FwdTcArgs {
    ProducerArgs pa;
    const float* features;
    const float* bg;
    float* final_T;
    uint32_t* n_contrib;
    float* out_color;
    float* out_feature;
    float* out_depth;
    int vec_store;  // bit1: 256-bit stores legal (W % 8 == 0, 32-byte aligned planes)
};

__global__ void __launch_bounds__(kThreadsTc, 1) composite_fwd_tc_kernel(const FwdTcArgs args) {
    extern __shared__ unsigned char smem_dyn[];
    // SWIZZLE_128B atoms need 1024-byte alignment: align by hand (the launcher asks for 1 KB of slack)
    TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u));
    RingSlim& ring = sm.ring;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int W = args.pa.W, H = args.pa.H, C = args.pa.C;
    const size_t HW = (size_t)H * W;

    ring_init<0>(ring, kAlphaN + kConvN + 1, false);
    if (threadIdx.x == 32) {
        for (int i = 0; i < kOpStages; i++) {
            mbar_init(&sm.op_full[i], kAlphaN + 1);
            mbar_init(&sm.op_empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.tmem_full[i], 1);
            mbar_init(&sm.tmem_empty[i], kEpiN);
        }
        mbar_fence_init();
    }
    {   // operands start as zeros: a stale row multiplied by a zero weight must not be NaN
        float4* p = reinterpret_cast<float4*>(&sm.op[0]);
        for (int i = threadIdx.x; i < (int)(sizeof(OpStage) * kOpStages / 16); i += blockDim.x)
            p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        fence_async_smem();
    }
    if (warp == kMmaWarp) tmem_alloc_512(&sm.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == kProducerWarp) {
        // ==================================================================== producer
        producer_loop<0, false, false, RingSlim>(ring, args.pa);
    } else if (warp == kMmaWarp) {
        // ==================================================================== MMA issuer
        constexpr uint32_t kIdesc = umma_idesc_tf32(128, 256, 1, 1);
        int s = 0, os = 0, buf = 0;
        uint32_t parity = 0, op_round = 0, tile_seq = 0;
        bool fresh = true;
        for (;;) {
            mbar_wait(&ring.full[s], parity);
            const Stage<0>& st = ring.stage[s];
            const uint32_t n = st.n, last = st.last, first = st.first;
            const int work = st.work;
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.empty[s]);
            if (first || work < 0) {
                buf = (int)(tile_seq & 1u);
                mbar_wait(&sm.tmem_empty[buf], ((tile_seq >> 1) & 1u) ^ 1u);
                if (lane == 0) {
                    sm.tile_work[buf] = work;
                    __threadfence_block();
                }
                tile_seq++;
                fresh = true;
            }
            if (work < 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.tmem_full[buf]);
                break;
            }
            const int nh = n > (uint32_t)kKS ? 2 : 1;
            for (int h = 0; h < nh; h++) {
                const int cnt = max(0, min(kKS, (int)n - h * kKS));
                const int ksteps = cnt > 8 ? 2 : 1;
                mbar_wait(&sm.op_full[os], op_round & 1u);
                tc_fence_after();
                if (lane == 0) {
                    const OpStage& op = sm.op[os];
                    const uint32_t fh = smem_u32(&op.Fhi[0][0][0]), fl = smem_u32(&op.Flo[0][0][0]);
                    const uint32_t wh = smem_u32(&op.Whi[0][0][0]), wl = smem_u32(&op.Wlo[0][0][0]);
                    const uint32_t d = tmem + (uint32_t)buf * 256u;
                    for (int g = 0; g < ksteps; g++) {
                        // MN-major SWIZZLE_128B_BASE32B: LBO = next 32-element block along M/N (16 rows x 128 B),
                        // SBO = next 4 k (512 B); the 8 instances of this step are the two atoms at g * 1024
                        const uint64_t a_hi = umma_desc(fh + g * 1024, kKS * 128, 512, kUmmaSw128Base32);
                        const uint64_t a_lo = umma_desc(fl + g * 1024, kKS * 128, 512, kUmmaSw128Base32);
                        const uint64_t b_hi = umma_desc(wh + g * 1024, kKS * 128, 512, kUmmaSw128Base32);
                        const uint64_t b_lo = umma_desc(wl + g * 1024, kKS * 128, 512, kUmmaSw128Base32);
                        umma_tf32_ss(d, a_hi, b_hi, kIdesc, (fresh && g == 0) ? 0u : 1u);
                        umma_tf32_ss(d, a_hi, b_lo, kIdesc, 1u);
                        umma_tf32_ss(d, a_lo, b_hi, kIdesc, 1u);
                    }
                    umma_commit(&sm.op_empty[os]);
                    if (last && h == nh - 1) umma_commit(&sm.tmem_full[buf]);
                }
                fresh = false;
                __syncwarp();
                if (++os == kOpStages) { os = 0; op_round++; }
            }
            if (++s == kStages) { s = 0; parity ^= 1; }
        }
    } else if (warp >= kAlpha0 && warp < kAlpha0 + kAlphaN) {
        // ==================================================================== alpha warps
        const int b = warp - kAlpha0;  // 8x4 pixel block of the tile; lane = pixel, row-major inside the block
        const int lx = lane & 7, ly = lane >> 3;
        int s = 0, os = 0, chunk = 0;
        uint32_t parity = 0, op_round = 0;
        float T = 1.f, Cr = 0.f, Cg = 0.f, Cb = 0.f, Dp = 0.f, pxf = 0.f, pyf = 0.f, fbx0 = 0.f, fby0 = 0.f;
        uint32_t last_contrib = 0;
        int px = 0, py = 0;
        bool done = true, inside = false, blk_done = true;
        for (;;) {
            mbar_wait(&ring.full[s], parity);
            Stage<0>& st = ring.stage[s];
            const uint32_t n = st.n, last = st.last, first = st.first;
            const int work = st.work;
            if (work < 0) break;
            if (first) {
                const int tile = work / args.pa.chunks;
                chunk = work - tile * args.pa.chunks;
                const int tile_x = tile % args.pa.tiles_x, tile_y = tile / args.pa.tiles_x;
                const int bx0 = tile_x * 16 + (b & 1) * 8, by0 = tile_y * 16 + (b >> 1) * 4;
                px = bx0 + lx;
                py = by0 + ly;
                inside = px < W && py < H;
                pxf = (float)px; pyf = (float)py; fbx0 = (float)bx0; fby0 = (float)by0;
                T = 1.f; Cr = Cg = Cb = Dp = 0.f;
                last_contrib = 0;
                done = !inside;
                blk_done = __all_sync(0xffffffffu, done);
                if (blk_done && lane == 0) atomicOr(&ring.done_mask[st.done_slot], 1u << b);
            }
            const int nh = n > (uint32_t)kKS ? 2 : 1;
            for (int h = 0; h < nh; h++) {
                mbar_wait(&sm.op_empty[os], (op_round & 1u) ^ 1u);
                OpStage& op = sm.op[os];
                float* whi = &op.Whi[b][0][0];
                float* wlo = &op.Wlo[b][0][0];
                {   // the block's 16 rows of both parts start as zeros (instances that miss the block stay zero)
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        reinterpret_cast<float4*>(whi)[i * 32 + lane] = z;
                        reinterpret_cast<float4*>(wlo)[i * 32 + lane] = z;
                    }
                }
                __syncwarp();
                if (!blk_done && n > 0) {
                    const int e0 = h * kKS;
                    bool hit = false;
                    if (lane < kKS && e0 + lane < (int)n)
                        hit = footprint_hits_rect(st.rec0[e0 + lane], st.rec1[e0 + lane], fbx0, fbx0 + 7.f, fby0, fby0 + 3.f);
                    uint32_t am = __ballot_sync(0xffffffffu, hit);
                    while (am) {
                        // up to 4 instances per trip, branch-free alpha evaluation (see composite_fwd.cu)
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
                            const float4 r0 = st.rec0[e0 + kk[u]];
                            const float4 r1 = st.rec1[e0 + kk[u]];
                            // same expression trees as reference forward.cu:340-351 (see common.cuh)
                            const float dx = r0.x - pxf, dy = r0.y - pyf;
                            const float power = -0.5f * (r1.x * dx * dx + r1.z * dy * dy) - r1.y * dx * dy;
                            const float av = fminf(0.99f, r1.w * expf(power));
                            al[u] = (vk[u] && !(power > 0.0f) && !(av < 1.0f / 255.0f)) ? av : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (!vk[u]) break;  // warp-uniform, only in the last trip
                            const float4 r2 = st.rec2[e0 + kk[u]];
                            const uint32_t lp = st.listpos[e0 + kk[u]];
                            const float alpha = al[u];
                            const float test_T = T * (1 - alpha);
                            const bool act = !done && alpha > 0.f;
                            const bool stop = act && (test_T < 0.0001f);  // reference: done = true, not blended
                            const bool blend = act && !stop;
                            done = done || stop;
                            const float wgt = blend ? alpha * T : 0.f;
                            const float nCr = Cr + r2.x * alpha * T;  // reference forward.cu:362-368
                            const float nCg = Cg + r2.y * alpha * T;
                            const float nCb = Cb + r2.z * alpha * T;
                            const float nDp = Dp + r2.w * (alpha * T);
                            Cr = blend ? nCr : Cr;
                            Cg = blend ? nCg : Cg;
                            Cb = blend ? nCb : Cb;
                            Dp = blend ? nDp : Dp;
                            T = blend ? test_T : T;
                            last_contrib = blend ? lp : last_contrib;
                            if (__any_sync(0xffffffffu, blend)) {
                                const float hi = tf32_hi(wgt);
                                const int o = sw32b_idx(kk[u], lane);
                                whi[o] = hi;
                                wlo[o] = wgt - hi;
                            }
                        }
                    }
                    if (__all_sync(0xffffffffu, done)) {
                        blk_done = true;
                        if (lane == 0) atomicOr(&ring.done_mask[st.done_slot], 1u << b);
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.op_full[os]);
                if (++os == kOpStages) { os = 0; op_round++; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.empty[s]);
            if (last && chunk == 0 && inside) {
                const size_t pix = (size_t)py * W + px;
                args.final_T[pix] = T;
                args.n_contrib[pix] = last_contrib;
                args.out_color[pix] = Cr + T * args.bg[0];  // reference forward.cu:389
                args.out_color[HW + pix] = Cg + T * args.bg[1];
                args.out_color[2 * HW + pix] = Cb + T * args.bg[2];
                args.out_depth[pix] = Dp;
            }
            if (++s == kStages) { s = 0; parity ^= 1; }
        }
    } else if (warp >= kConv0 && warp < kConv0 + kConvN) {
        // ==================================================================== convert warps
        const int cw = warp - kConv0;
        int s = 0;
        uint32_t parity = 0, seq = 0;  // seq: operand stages since the start (this warp serves seq % kConvN == cw)
        for (;;) {
            mbar_wait(&ring.full[s], parity);
            const Stage<0>& st = ring.stage[s];
            const uint32_t n = st.n;
            const int work = st.work;
            if (work < 0) break;
            const int chunk_off = (work % args.pa.chunks) * 128;
            const int row_floats = min(128, C - chunk_off);
            const int nh = n > (uint32_t)kKS ? 2 : 1;
            for (int h = 0; h < nh; h++, seq++) {
                if ((int)(seq % kConvN) != cw) continue;
                const int cnt = max(0, min(kKS, (int)n - h * kKS));
                float4 v[kKS];
#pragma unroll
                for (int r = 0; r < kKS; r++) {
                    v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < cnt && 4 * lane < row_floats)
                        v[r] = ld_nc_f4(args.features + (size_t)st.gid[h * kKS + r] * C + chunk_off + 4 * lane);
                }
                const int os = (int)(seq % kOpStages);
                mbar_wait(&sm.op_empty[os], ((seq / kOpStages) & 1u) ^ 1u);
                OpStage& op = sm.op[os];
                float* fhi = &op.Fhi[lane >> 3][0][0];
                float* flo = &op.Flo[lane >> 3][0][0];
#pragma unroll
                for (int r = 0; r < kKS; r++) {
                    if (r < cnt) {
                        const float4 hi = make_float4(tf32_hi(v[r].x), tf32_hi(v[r].y), tf32_hi(v[r].z), tf32_hi(v[r].w));
                        const float4 lo = make_float4(v[r].x - hi.x, v[r].y - hi.y, v[r].z - hi.z, v[r].w - hi.w);
                        // 16-byte piece (lane & 7) of the row: 32-byte chunk ((lane & 7) >> 1) ^ (r & 3), half (lane & 1)
                        const int o = r * 32 + (((((lane & 7) >> 1) ^ (r & 3)) << 1) | (lane & 1)) * 4;
                        *reinterpret_cast<float4*>(fhi + o) = hi;
                        *reinterpret_cast<float4*>(flo + o) = lo;
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.op_full[os]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.empty[s]);
            if (++s == kStages) { s = 0; parity ^= 1; }
        }
    } else if (warp >= kEpi0) {
        // ==================================================================== epilogue warps
        const int ew = warp - kEpi0;
        for (uint32_t seq = 0;; seq++) {
            const int buf = (int)(seq & 1u);
            mbar_wait(&sm.tmem_full[buf], (seq >> 1) & 1u);
            tc_fence_after();
            const int work = *reinterpret_cast<volatile int32_t*>(&sm.tile_work[buf]);
            if (work < 0) break;
            const int tile = work / args.pa.chunks, chunk = work - tile * args.pa.chunks;
            const int tile_x = tile % args.pa.tiles_x, tile_y = tile / args.pa.tiles_x;
            const int ch = chunk * 128 + 32 * ew + lane;
            float* plane = args.out_feature + (size_t)min(ch, C - 1) * HW;
#pragma unroll 1
            for (int j = 0; j < kBlocksPerTile; j++) {
                uint32_t r[32];
                tmem_ld_x32(tmem + ((uint32_t)(32 * ew) << 16) + (uint32_t)(buf * 256 + j * 32), r);
                tmem_ld_wait();
                const int bx0 = tile_x * 16 + (j & 1) * 8, by0 = tile_y * 16 + (j >> 1) * 4;
                if (ch < C) {
#pragma unroll
                    for (int y = 0; y < 4; y++) {
                        const int yy = by0 + y;
                        if (yy >= H) continue;
                        float* row = plane + (size_t)yy * W + bx0;
                        if ((args.vec_store & 2) && bx0 + 8 <= W) {
                            st_na_f8(row,
                                     make_float4(__uint_as_float(r[8 * y]), __uint_as_float(r[8 * y + 1]),
                                                 __uint_as_float(r[8 * y + 2]), __uint_as_float(r[8 * y + 3])),
                                     make_float4(__uint_as_float(r[8 * y + 4]), __uint_as_float(r[8 * y + 5]),
                                                 __uint_as_float(r[8 * y + 6]), __uint_as_float(r[8 * y + 7])));
                        } else {
#pragma unroll
                            for (int x = 0; x < 8; x++)
                                if (bx0 + x < W) row[x] = __uint_as_float(r[8 * y + x]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tmem_empty[buf]);
        }
    }

    // ======================================================================== teardown
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc_512(tmem);
    }
}

}  // namespace

// Same contract as launch_composite_fwd (composite_fwd.cu).  Preconditions checked by the caller (api.cu): C % 4 == 0 and a
// 16-byte aligned feature matrix; every other shape takes the fp32-pipe kernel.
cudaError_t launch_composite_fwd_tc(const ViewParams& vp, const uint2* ranges, const uint32_t* point_list,
                                    const SplatRec* rec, const float* features, const float* bg, float* final_T,
                                    uint32_t* n_contrib, float* out_color, float* out_feature, float* out_depth,
                                    int* work_counter, cudaStream_t s) {
    const size_t smem = sizeof(TcSmem) + 1024;
    static std::atomic<int> sms_of_device[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (sms_of_device[dev].load() == 0) {
        cudaError_t e = cudaFuncSetAttribute(composite_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms_of_device[dev].store(n > 0 ? n : 148);
    }
    FwdTcArgs a;
    a.pa.ranges = ranges; a.pa.point_list = point_list; a.pa.rec = rec;
    a.pa.features = nullptr;  // the record ring carries no feature rows: the convert warps fetch them
    a.pa.n_contrib = nullptr;
    a.pa.work_counter = work_counter;
    a.pa.W = vp.W; a.pa.H = vp.H; a.pa.C = vp.C;
    a.pa.tiles_x = (int)vp.grid_x;
    a.pa.num_tiles = (int)(vp.grid_x * vp.grid_y);
    a.pa.chunks = (vp.C + 127) / 128;
    a.pa.use_bulk = 0;
    a.features = features;
    a.bg = bg; a.final_T = final_T; a.n_contrib = n_contrib;
    a.out_color = out_color; a.out_feature = out_feature; a.out_depth = out_depth;
    a.vec_store = (vp.W % 8 == 0 && (reinterpret_cast<uintptr_t>(out_feature) & 31) == 0) ? 2 : 0;
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    const int grid = min(a.pa.num_tiles * a.pa.chunks, sms_of_device[dev].load());
    composite_fwd_tc_kernel<<<grid, kThreadsTc, smem, s>>>(a);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace f3dgs
