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
#include <cstdio>
#include <cstdlib>

#include "composite_common.cuh"
#include "tc_common.cuh"

#ifndef F3DGS_TIMING_BUILD
#define F3DGS_TIMING_BUILD 0   // 1 (tools/build_variants.sh "timing"): per-role cycle counters, printed when F3DGS_TIMING is set
#endif

namespace f3dgs {

namespace {

// Two configurations (NCTA = CTAs per SM).  The kernel is bound by the alpha warps' dependent chains (8 warps per CTA), so
// the default is TWO CTAs per SM: 16 alpha warps per SM hide each other's latencies; each CTA then owns half of the
// tensor memory (one accumulator buffer: its epilogue overlaps the other CTA's MMAs) and half of the shared memory.
//   NCTA = 1: 20 warps (producer, MMA, 2 idle, 8 alpha, 4 convert, 4 epilogue), 4 operand stages, two accumulator buffers
//   NCTA = 2: 16 warps (producer, MMA, 2 convert, 8 alpha, 4 epilogue),         2 operand stages, one accumulator buffer
#ifndef F3DGS_TC_CTAS
#define F3DGS_TC_CTAS 1
#endif
constexpr int kNCta = F3DGS_TC_CTAS;
constexpr int kKS = 16;                          // instances per operand stage (two K=8 MMA steps)
constexpr int kOpStages = kNCta == 1 ? 4 : 2;    // operand ring depth
constexpr int kAccBufs = kNCta == 1 ? 2 : 1;     // accumulator buffers of 256 columns
constexpr int kTmemCols = 256 * kAccBufs;
constexpr int kMmaWarp = 1;
constexpr int kAlpha0 = 4, kAlphaN = 8;
constexpr int kConv0 = kNCta == 1 ? 12 : 2, kConvN = kNCta == 1 ? 4 : 2;
constexpr int kEpi0 = kNCta == 1 ? 16 : 12, kEpiN = 4;   // kEpi0 % 4 == 0: epilogue warp e reads TMEM lanes 32e .. 32e+31
constexpr int kThreadsTc = (kEpi0 + kEpiN) * 32;
#ifndef F3DGS_TC_CONV_BATCH
#define F3DGS_TC_CONV_BATCH (F3DGS_TC_CTAS == 1 ? 16 : 8)
#endif
constexpr int kConvBatch = F3DGS_TC_CONV_BATCH;  // feature rows a convert warp keeps in flight
// diagnostic builds only (WRONG RESULTS): drop one role's work to see what bounds the kernel (tools/build_variants.sh)
#ifndef F3DGS_TC_DIAG
#define F3DGS_TC_DIAG 0   // 1: no MMAs, 2: alpha warps skip the evaluation, 4: no zero fill of the weight rows, 8: no feature loads
#endif
#ifndef F3DGS_TC_BLOCK_EXACT
#define F3DGS_TC_BLOCK_EXACT 0   // 1: exact ellipse test at block level too (the producer's tile-level test follows F3DGS_EXACT_CULL)
#endif

struct alignas(1024) OpStage {
    float Fhi[4][kKS][32];   // A operand: [channel block][instance][32 channels]
    float Flo[4][kKS][32];
    float Whi[8][kKS][32];   // B operand: [pixel block][instance][32 pixels]
    float Wlo[8][kKS][32];
};
static_assert(sizeof(OpStage) == 48 * 1024, "operand stage is 48 KB");

struct alignas(1024) TcSmem {
    OpStage op[kOpStages];
    RingRec ring;                        // record ring (composite_common.cuh)
    uint64_t op_full[kOpStages];         // 8 alpha warps + 1 convert warp have written the stage
    uint64_t op_empty[kOpStages];        // the stage's MMAs have completed (tcgen05.commit)
    uint64_t tmem_full[kAccBufs];        // a tile's accumulators are complete
    uint64_t tmem_empty[kAccBufs];       // the epilogue has read them
    int32_t tile_work[kAccBufs];
    uint32_t tmem_base;
};

static_assert(sizeof(TcSmem) + 1024 <= (kNCta == 1 ? 227 * 1024 : 112 * 1024), "shared memory budget per CTA");

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
    long long* dbg;  // timing builds only: [cta][warp][8] cycle counters, else nullptr
};
constexpr bool kTimingTc = F3DGS_TIMING_BUILD != 0;
#define TTICK() ((kTimingTc && args.dbg) ? clock64() : 0ll)
#define TWAIT(acc, bar, par) do { const long long t_ = TTICK(); mbar_wait(bar, par); acc += TTICK() - t_; } while (0)
#define TWAITS(acc, bar, par, ns) do { const long long t_ = TTICK(); mbar_wait_sleep(bar, par, ns); acc += TTICK() - t_; } while (0)
#define TDUMP(a0, a1, a2, a3, a4, a5)                                                                   \
    do {                                                                                                \
        if (kTimingTc && args.dbg && lane == 0) {                                                       \
            long long* d_ = args.dbg + ((size_t)blockIdx.x * 32 + warp) * 8;                            \
            d_[0] = (a0); d_[1] = (a1); d_[2] = (a2); d_[3] = (a3); d_[4] = (a4); d_[5] = (a5);         \
        }                                                                                               \
    } while (0)

__global__ void __launch_bounds__(kThreadsTc, kNCta) composite_fwd_tc_kernel(const FwdTcArgs args) {
    extern __shared__ unsigned char smem_dyn[];
    // SWIZZLE_128B atoms need 1024-byte alignment: align by hand (the launcher asks for 1 KB of slack)
    TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u));
    RingRec& ring = sm.ring;
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
        for (int i = 0; i < kAccBufs; i++) {
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
    if (warp == kMmaWarp) tmem_alloc<kTmemCols>(&sm.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == kProducerWarp) {
        // ==================================================================== producer
        const long long t_tot = TTICK();
        producer_loop<0, false, false, RingRec>(ring, args.pa);
        TDUMP(TTICK() - t_tot, 0, 0, 0, 0, 0);
    } else if (warp == kMmaWarp) {
        // ==================================================================== MMA issuer
        constexpr uint32_t kIdesc = umma_idesc_tf32(128, 256, 1, 1);
        int s = 0, os = 0, buf = 0;
        uint32_t parity = 0, op_round = 0, tile_seq = 0;
        bool fresh = true;
        long long t_tot = TTICK(), w_full = 0, w_tmem = 0, w_op = 0, n_st = 0;
        for (;;) {
            TWAIT(w_full, &ring.full[s], parity);
            const Stage<0>& st = ring.stage[s];
            const uint32_t n = st.n, last = st.last, first = st.first;
            const int work = st.work;
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.empty[s]);
            if (first || work < 0) {
                buf = (int)(tile_seq % kAccBufs);
                TWAITS(w_tmem, &sm.tmem_empty[buf], ((tile_seq / kAccBufs) & 1u) ^ 1u, 64);
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
                TWAITS(w_op, &sm.op_full[os], op_round & 1u, 32);
                n_st++;
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
                        if (!(F3DGS_TC_DIAG & 1)) {
                            umma_tf32_ss(d, a_hi, b_hi, kIdesc, (fresh && g == 0) ? 0u : 1u);
                            umma_tf32_ss(d, a_hi, b_lo, kIdesc, 1u);
                            umma_tf32_ss(d, a_lo, b_hi, kIdesc, 1u);
                        }
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
        TDUMP(TTICK() - t_tot, w_full, w_tmem, w_op, n_st, 0);
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
        long long t_tot = TTICK(), w_full = 0, w_op = 0, n_st = 0, n_hit = 0;
        for (;;) {
            TWAIT(w_full, &ring.full[s], parity);
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
                TWAIT(w_op, &sm.op_empty[os], (op_round & 1u) ^ 1u);
                n_st++;
                OpStage& op = sm.op[os];
                float* whi = &op.Whi[b][0][0];
                float* wlo = &op.Wlo[b][0][0];
                if (!(F3DGS_TC_DIAG & 4)) {   // the block's 16 rows of both parts start as zeros (instances that miss the block stay zero)
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        reinterpret_cast<float4*>(whi)[i * 32 + lane] = z;
                        reinterpret_cast<float4*>(wlo)[i * 32 + lane] = z;
                    }
                }
                __syncwarp();
                if (!blk_done && n > 0 && !(F3DGS_TC_DIAG & 2)) {
                    const int e0 = h * kKS;
                    bool hit = false;
                    if (lane < kKS && e0 + lane < (int)n) {
#if F3DGS_TC_BLOCK_EXACT
                        hit = footprint_hits_rect(st.rec0[e0 + lane], st.rec1[e0 + lane], fbx0, fbx0 + 7.f, fby0, fby0 + 3.f);
#else
                        const float4 r0 = st.rec0[e0 + lane];
                        hit = (r0.x + r0.z >= fbx0) && (r0.x - r0.z <= fbx0 + 7.f) && (r0.y + r0.w >= fby0) &&
                              (r0.y - r0.w <= fby0 + 3.f);
#endif
                    }
                    uint32_t am = __ballot_sync(0xffffffffu, hit);
                    if (kTimingTc) n_hit += __popc(am);
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
                        // The T recurrence of the (up to) four instances, written WITHOUT control flow: an empty slot has
                        // alpha = 0 and changes nothing.  The only true dependency between consecutive instances is
                        // T -> test_T -> T (and `done`); everything else (colour / depth accumulation, the hi / lo split,
                        // the stores) then overlaps across the four.  With the `break` / `if (any blend)` branches of
                        // the first version the warp ran one dependent chain per instance: 408 cycles per hit, 21% of the
                        // cycles issuing (ncu source counters), and bounded the whole kernel.
                        float4 r2v[4];
                        uint32_t lpv[4];
                        float wg[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            r2v[u] = st.rec2[e0 + kk[u]];
                            lpv[u] = st.listpos[e0 + kk[u]];
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const float4 r2 = r2v[u];
                            const float alpha = al[u];
                            const float test_T = T * (1 - alpha);
                            const bool act = !done && alpha > 0.f;
                            const bool stop = act && (test_T < 0.0001f);  // reference: done = true, not blended
                            const bool blend = act && !stop;
                            done = done || stop;
                            wg[u] = blend ? alpha * T : 0.f;
                            const float nCr = Cr + r2.x * alpha * T;  // reference forward.cu:362-368
                            const float nCg = Cg + r2.y * alpha * T;
                            const float nCb = Cb + r2.z * alpha * T;
                            const float nDp = Dp + r2.w * (alpha * T);
                            Cr = blend ? nCr : Cr;
                            Cg = blend ? nCg : Cg;
                            Cb = blend ? nCb : Cb;
                            Dp = blend ? nDp : Dp;
                            T = blend ? test_T : T;
                            last_contrib = blend ? lpv[u] : last_contrib;
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (vk[u]) {  // warp-uniform
                                const float hi = tf32_hi(wg[u]);
                                const int o = sw32b_idx(kk[u], lane);
                                whi[o] = hi;
                                wlo[o] = wg[u] - hi;
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
        TDUMP(TTICK() - t_tot, w_full, w_op, n_st, n_hit, 0);
    } else if (warp >= kConv0 && warp < kConv0 + kConvN) {
        // ==================================================================== convert warps
        const int cw = warp - kConv0;
        int s = 0;
        uint32_t parity = 0, seq = 0;  // seq: operand stages since the start (this warp serves seq % kConvN == cw)
        long long t_tot = TTICK(), w_full = 0, w_op = 0, n_st = 0;
        for (;;) {
            TWAITS(w_full, &ring.full[s], parity, 64);
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
                const int os = (int)(seq % kOpStages);
                OpStage& op = sm.op[os];
                float* fhi = &op.Fhi[lane >> 3][0][0];
                float* flo = &op.Flo[lane >> 3][0][0];
                bool waited = false;
#pragma unroll
                for (int r0 = 0; r0 < kKS; r0 += kConvBatch) {
                    float4 v[kConvBatch];
#pragma unroll
                    for (int i = 0; i < kConvBatch; i++) {
                        const int r = r0 + i;
                        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (r < cnt && 4 * lane < row_floats && !(F3DGS_TC_DIAG & 8))
                            v[i] = ld_nc_f4(args.features + (size_t)st.gid[h * kKS + r] * C + chunk_off + 4 * lane);
                    }
                    if (!waited) {  // the loads are in flight while the stage's previous MMAs drain
                        TWAITS(w_op, &sm.op_empty[os], ((seq / kOpStages) & 1u) ^ 1u, 128);
                        n_st++;
                        waited = true;
                    }
#pragma unroll
                    for (int i = 0; i < kConvBatch; i++) {
                        const int r = r0 + i;
                        if (r < cnt) {
                            const float4 hi = make_float4(tf32_hi(v[i].x), tf32_hi(v[i].y), tf32_hi(v[i].z), tf32_hi(v[i].w));
                            const float4 lo = make_float4(v[i].x - hi.x, v[i].y - hi.y, v[i].z - hi.z, v[i].w - hi.w);
                            // 16-byte piece (lane & 7) of the row: 32-byte chunk ((lane & 7) >> 1) ^ (r & 3), half (lane & 1)
                            const int o = r * 32 + (((((lane & 7) >> 1) ^ (r & 3)) << 1) | (lane & 1)) * 4;
                            *reinterpret_cast<float4*>(fhi + o) = hi;
                            *reinterpret_cast<float4*>(flo + o) = lo;
                        }
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
        TDUMP(TTICK() - t_tot, w_full, w_op, n_st, 0, 0);
    } else if (warp >= kEpi0) {
        // ==================================================================== epilogue warps
        const int ew = warp - kEpi0;
        long long t_tot = TTICK(), w_full = 0, n_tiles = 0;
        for (uint32_t seq = 0;; seq++) {
            const int buf = (int)(seq % kAccBufs);
            TWAITS(w_full, &sm.tmem_full[buf], (seq / kAccBufs) & 1u, 512);
            n_tiles++;
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
        TDUMP(TTICK() - t_tot, w_full, n_tiles, 0, 0, 0);
    }

    // ======================================================================== teardown
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem);
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
    const int grid = min(a.pa.num_tiles * a.pa.chunks, kNCta * sms_of_device[dev].load());
    static long long* dbg = nullptr;
    const bool timing = kTimingTc && getenv("F3DGS_TIMING") != nullptr;
    if (timing && !dbg) cudaMalloc(&dbg, 512 * 32 * 8 * sizeof(long long));
    a.dbg = timing ? dbg : nullptr;
    if (timing) cudaMemsetAsync(dbg, 0, 512 * 32 * 8 * sizeof(long long), s);
    composite_fwd_tc_kernel<<<grid, kThreadsTc, smem, s>>>(a);
    g_launches++;
    if (timing) {
        static long long host[512 * 32 * 8];
        cudaMemcpyAsync(host, dbg, sizeof(host), cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        auto avg = [&](int w0, int nw, int i) {
            double t = 0;
            for (int c = 0; c < grid; c++) for (int w = w0; w < w0 + nw; w++) t += host[((size_t)c * 32 + w) * 8 + i];
            return t / ((double)grid * nw);
        };
        fprintf(stderr, "[f3dgs timing fwd_tc] per-warp mean cycles: producer %.0f | mma total %.0f wait_rec %.0f wait_tmem %.0f wait_opfull %.0f stages %.0f | "
                        "alpha total %.0f wait_rec %.0f wait_opempty %.0f stages %.0f hits %.0f | convert total %.0f wait_rec %.0f wait_opempty %.0f stages %.0f | "
                        "epilogue total %.0f wait_tmemfull %.0f tiles %.0f\n",
                avg(0, 1, 0), avg(kMmaWarp, 1, 0), avg(kMmaWarp, 1, 1), avg(kMmaWarp, 1, 2), avg(kMmaWarp, 1, 3), avg(kMmaWarp, 1, 4),
                avg(kAlpha0, kAlphaN, 0), avg(kAlpha0, kAlphaN, 1), avg(kAlpha0, kAlphaN, 2), avg(kAlpha0, kAlphaN, 3), avg(kAlpha0, kAlphaN, 4),
                avg(kConv0, kConvN, 0), avg(kConv0, kConvN, 1), avg(kConv0, kConvN, 2), avg(kConv0, kConvN, 3),
                avg(kEpi0, kEpiN, 0), avg(kEpi0, kEpiN, 1), avg(kEpi0, kEpiN, 2));
    }
    return cudaGetLastError();
}

}  // namespace f3dgs
