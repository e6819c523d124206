// Backward tile composite: walks each tile's instance list back to front and produces the
// per-Gaussian gradients of colour, feature, depth, opacity, conic and 2-D mean.
// Reference: backward.cu:407-620 (renderCUDA<3>), semantics restated in SURVEY.md A.5 / D.1.
//
// The reference issues (C+10) same-address global atomics per blended (pixel, Gaussian) pair.
// Here every pair contributes to a reduction inside the warp that owns the 8x4 pixel block first:
//   geometry/colour/depth terms (10 scalars per pair): lanes (= pixels) park their terms in a
//       [value][instance slot][pixel] shared tile; when 8 instances are parked the tile is summed
//       row-wise (one lane per row, conflict-free padded rows) and each row sum becomes ONE
//       red.global.add -> 32x fewer atomics, none of them contended inside the warp;
//   feature terms: lane = float4 of channels keeps dL/dfeature_map of all 32 pixels x 4 channels
//       in registers for the whole kernel (the upstream gradient is read from HBM exactly once);
//       per blended instance  g[4] = sum_pixels w[pixel] * dO[pixel][4]  is 2x2-quad sparse FMAs
//       fed by broadcast LDS.128 of the blend weights, then one red.global.add.v4.f32 per lane:
//       a fully coalesced 512-byte vector reduction per (warp, instance) at C = 128.
// As in the reference, the feature loss does not feed dL/dalpha (backward.cu:575 is disabled).
// No features are read at all: the feature gradient needs only w = alpha*T and dL/dout.
#include "composite_common.cuh"

namespace f3dgs {

constexpr int kBwdStages = 4;
constexpr int kRedSlots = 8;
constexpr int kRedVals = 10;
constexpr int kRedRows = kRedSlots * kRedVals;

struct BwdSmem {
    Ring<0, kBwdStages> ring;
    float w[kConsumerWarps][kStageEntries][32];
    float red[kConsumerWarps][kRedRows][33];
    uint32_t red_gid[kConsumerWarps][kRedSlots];
};

struct BwdOut {
    float* dL_dmean2D;   // [P,3]
    float* dL_dconic;    // [P,4]
    float* dL_dopacity;  // [P]
    float* dL_dcolor;    // [P,3]
    float* dL_dfeature;  // [P,C]
    float* dL_dz;        // [P]
};

template <int CH>
__global__ void __launch_bounds__(kBlockThreads, 1)
composite_bwd_kernel(int W, int H, int C, const uint2* __restrict__ ranges,
                     const uint32_t* __restrict__ point_list, const SplatRec* __restrict__ rec,
                     const float* __restrict__ bg, const float* __restrict__ final_T,
                     const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                     const float* __restrict__ dL_dfeat_pix, const float* __restrict__ dL_ddepth, BwdOut out,
                     int vec_io) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
    Ring<0, kBwdStages>& ring = sm.ring;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y, chunk = blockIdx.z;
    const int chunk_off = chunk * CH;
    const uint2 range = ranges[tile_y * gridDim.x + tile_x];
    const size_t HW = (size_t)H * W;

    ring_init(ring);
    __syncthreads();

    if (warp == kConsumerWarps) {
        // nothing behind the deepest last-contributor of the tile is ever used
        uint32_t tmax = 0;
        {
            const int yy = tile_y * 16 + (lane >> 1), xb = tile_x * 16 + (lane & 1) * 8;
            if (yy < H)
                for (int i = 0; i < 8; i++)
                    if (xb + i < W) tmax = max(tmax, n_contrib[(size_t)yy * W + xb + i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        }
        const uint32_t walk = min(range.y - range.x, tmax);
        const float tx0 = (float)(tile_x * 16), ty0 = (float)(tile_y * 16);
        producer_loop<0, kBwdStages, true>(ring, point_list, rec, nullptr, 0, 0, 0, false, range.x, range.y, walk,
                                           tx0, ty0, tx0 + 15.f, ty0 + 15.f);
        return;
    }

    // ------------------------------------------------------------------ consumer warps
    constexpr int LPR = CH > 0 ? CH / 4 : 32;
    constexpr int G = 32 / LPR;
    constexpr int NQ = 8 / G;
    const int grp = lane / LPR, cl = lane % LPR;

    const int bx0 = tile_x * 16 + (warp & 1) * 8, by0 = tile_y * 16 + (warp >> 1) * 4;
    const int px = bx0 + lane_px(lane), py = by0 + lane_py(lane);
    const bool inside = px < W && py < H;
    const size_t pix = inside ? (size_t)py * W + px : 0;
    const float pxf = (float)px, pyf = (float)py;
    const float fbx0 = (float)bx0, fby0 = (float)by0, fbx1 = (float)(bx0 + 7), fby1 = (float)(by0 + 3);

    const float T_final = inside ? final_T[pix] : 0.f;
    float T = T_final;
    const uint32_t last_contrib = inside ? n_contrib[pix] : 0u;
    float dLp[3] = {0.f, 0.f, 0.f}, dLd = 0.f;
    if (inside) {
        dLp[0] = dL_dpix[pix];
        dLp[1] = dL_dpix[HW + pix];
        dLp[2] = dL_dpix[2 * HW + pix];
        dLd = dL_ddepth[pix];
    }
    const float bg_dot = bg[0] * dLp[0] + bg[1] * dLp[1] + bg[2] * dLp[2];
    float accum_rec[3] = {0.f, 0.f, 0.f}, last_color[3] = {0.f, 0.f, 0.f};
    float last_alpha = 0.f, accum_depth = 0.f, last_depth = 0.f;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

    uint32_t wmax = last_contrib;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));

    // upstream feature gradient of the warp's 32 pixels x this lane's 4 channels, kept in registers
    float dO[NQ][4][4];
#pragma unroll
    for (int q = 0; q < NQ; q++)
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int c = 0; c < 4; c++) dO[q][i][c] = 0.f;
    if (CH > 0) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int ch = chunk_off + cl * 4 + c;
            if (ch >= C) continue;
            const float* plane = dL_dfeat_pix + (size_t)ch * HW;
            if (G == 1 && (vec_io & 1)) {
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    const int yy = by0 + y;
                    if (yy >= H) continue;
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        const int xx = bx0 + half * 4;
                        if (xx >= W) continue;
                        const int qa = (y >> 1) * 4 + half * 2, i0 = (y & 1) * 2;
                        const float4 v = ld_nc_f4(plane + (size_t)yy * W + xx);
                        dO[qa % NQ][i0][c] = v.x;
                        dO[qa % NQ][i0 + 1][c] = v.y;
                        dO[(qa + 1) % NQ][i0][c] = v.z;
                        dO[(qa + 1) % NQ][i0 + 1][c] = v.w;
                    }
                }
            } else {
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) {
                    const int q = qi * G + grp;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int xx = bx0 + (q & 3) * 2 + (i & 1), yy = by0 + (q >> 2) * 2 + (i >> 1);
                        if (xx < W && yy < H) dO[qi][i][c] = __ldg(plane + (size_t)yy * W + xx);
                    }
                }
            }
        }
    }

    float(*wbuf)[32] = sm.w[warp];
    float(*red)[33] = sm.red[warp];
    uint32_t* red_gid = sm.red_gid[warp];
    uint32_t nslots = 0;  // warp-uniform
    const bool do_geom = (chunk == 0);

    auto flush = [&]() {
        __syncwarp();
        for (int r = lane; r < kRedRows; r += 32) {
            const int slot = r % kRedSlots, v = r / kRedSlots;
            if (slot < (int)nslots) {
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    s0 += red[r][j];
                    s1 += red[r][j + 1];
                    s2 += red[r][j + 2];
                    s3 += red[r][j + 3];
                }
                const float sum = (s0 + s1) + (s2 + s3);
                const uint32_t gid = red_gid[slot];
                float* dst;
                switch (v) {
                    case 0: dst = out.dL_dmean2D + 3 * (size_t)gid; break;
                    case 1: dst = out.dL_dmean2D + 3 * (size_t)gid + 1; break;
                    case 2: dst = out.dL_dconic + 4 * (size_t)gid; break;
                    case 3: dst = out.dL_dconic + 4 * (size_t)gid + 1; break;
                    case 4: dst = out.dL_dconic + 4 * (size_t)gid + 3; break;
                    case 5: dst = out.dL_dopacity + gid; break;
                    case 6: dst = out.dL_dz + gid; break;
                    default: dst = out.dL_dcolor + 3 * (size_t)gid + (v - 7); break;
                }
                red_add_f1(dst, sum);
            }
        }
        __syncwarp();
        nslots = 0;
    };

    int s = 0;
    uint32_t parity = 0;
    while (true) {
        mbar_wait(&ring.full[s], parity);
        Stage<0>& st = ring.stage[s];
        const uint32_t n = st.n;
        const uint32_t last = st.last;
        if (n > 0 && wmax > 0) {
            bool hit = false;
            if (lane < n) {
                const float4 r0 = st.rec0[lane];
                hit = (st.listpos[lane] <= wmax) && (r0.x + r0.z >= fbx0) && (r0.x - r0.z <= fbx1) &&
                      (r0.y + r0.w >= fby0) && (r0.y - r0.w <= fby1);
            }
            uint32_t am = __ballot_sync(0xffffffffu, hit);
            uint32_t mypm = 0;
            while (am) {
                int kk[4];
                float al[4], Gv[4], dxv[4], dyv[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    kk[u] = am ? (__ffs(am) - 1) : -1;
                    am &= am - 1;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    al[u] = 0.f; Gv[u] = 0.f; dxv[u] = 0.f; dyv[u] = 0.f;
                    if (kk[u] >= 0) {
                        const float4 r0 = st.rec0[kk[u]];
                        const float4 r1 = st.rec1[kk[u]];
                        // same expression trees as reference backward.cu:525-535
                        const float dx = r0.x - pxf, dy = r0.y - pyf;
                        const float power = -0.5f * (r1.x * dx * dx + r1.z * dy * dy) - r1.y * dx * dy;
                        if (!(power > 0.0f)) {
                            const float Gs = expf(power);
                            const float a = fminf(0.99f, r1.w * Gs);
                            if (!(a < 1.0f / 255.0f)) {
                                al[u] = a; Gv[u] = Gs; dxv[u] = dx; dyv[u] = dy;
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (kk[u] < 0) break;  // warp-uniform
                    const int k = kk[u];
                    const float alpha = al[u];
                    const bool contrib = inside && alpha > 0.f && st.listpos[k] <= last_contrib;
                    float wgt = 0.f;
                    float v[kRedVals];
#pragma unroll
                    for (int i = 0; i < kRedVals; i++) v[i] = 0.f;
                    if (contrib) {
                        const float4 r1 = st.rec1[k];
                        const float4 r2 = st.rec2[k];
                        const float one_m_a = 1.f - alpha;
                        T = T / one_m_a;
                        wgt = alpha * T;
                        float dL_dalpha = 0.f;
                        const float col[3] = {r2.x, r2.y, r2.z};
#pragma unroll
                        for (int ch = 0; ch < 3; ch++) {
                            accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                            last_color[ch] = col[ch];
                            dL_dalpha += (col[ch] - accum_rec[ch]) * dLp[ch];
                            v[7 + ch] = wgt * dLp[ch];
                        }
                        accum_depth = last_alpha * last_depth + (1.f - last_alpha) * accum_depth;
                        last_depth = r2.w;
                        dL_dalpha += (r2.w - accum_depth) * dLd;
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final / one_m_a) * bg_dot;
                        const float Gs = Gv[u], dx = dxv[u], dy = dyv[u];
                        const float dL_dG = r1.w * dL_dalpha;
                        const float gdx = Gs * dx, gdy = Gs * dy;
                        const float dG_ddelx = -gdx * r1.x - gdy * r1.y;
                        const float dG_ddely = -gdy * r1.z - gdx * r1.y;
                        v[0] = dL_dG * dG_ddelx * ddelx_dx;
                        v[1] = dL_dG * dG_ddely * ddely_dy;
                        v[2] = -0.5f * gdx * dx * dL_dG;
                        v[3] = -0.5f * gdx * dy * dL_dG;
                        v[4] = -0.5f * gdy * dy * dL_dG;
                        v[5] = Gs * dL_dalpha;
                        v[6] = wgt * dLd;
                    }
                    const uint32_t pm = __ballot_sync(0xffffffffu, contrib);
                    if (pm) {
                        if (CH > 0) {
                            wbuf[k][lane] = wgt;
                            if (lane == k) mypm = pm;
                        }
                        if (do_geom) {
#pragma unroll
                            for (int i = 0; i < kRedVals; i++) red[i * kRedSlots + nslots][lane] = v[i];
                            if (lane == 0) red_gid[nslots] = st.gid[k];
                            nslots++;
                            if (nslots == kRedSlots) flush();
                        }
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
                    float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
                    for (int qi = 0; qi < NQ; qi++) {
                        const int q = qi * G + grp;
                        if ((pm >> (4 * q)) & 0xFu) {
                            const float4 w4 = *reinterpret_cast<const float4*>(&wbuf[k][4 * q]);
                            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                g0 = fmaf(wv[i], dO[qi][i][0], g0);
                                g1 = fmaf(wv[i], dO[qi][i][1], g1);
                                g2 = fmaf(wv[i], dO[qi][i][2], g2);
                                g3 = fmaf(wv[i], dO[qi][i][3], g3);
                            }
                        }
                    }
#pragma unroll
                    for (int o = LPR; o < 32; o <<= 1) {
                        g0 += __shfl_xor_sync(0xffffffffu, g0, o);
                        g1 += __shfl_xor_sync(0xffffffffu, g1, o);
                        g2 += __shfl_xor_sync(0xffffffffu, g2, o);
                        g3 += __shfl_xor_sync(0xffffffffu, g3, o);
                    }
                    const int ch = chunk_off + cl * 4;
                    if (grp == 0 && ch < C) {
                        float* dst = out.dL_dfeature + (size_t)st.gid[k] * C + ch;
                        if (vec_io & 2) {
                            red_add_f4(dst, make_float4(g0, g1, g2, g3));
                        } else {
                            red_add_f1(dst, g0);
                            if (ch + 1 < C) red_add_f1(dst + 1, g1);
                            if (ch + 2 < C) red_add_f1(dst + 2, g2);
                            if (ch + 3 < C) red_add_f1(dst + 3, g3);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring.empty[s]);
        if (last) break;
        s++;
        if (s == kBwdStages) {
            s = 0;
            parity ^= 1;
        }
    }
    if (do_geom && nslots > 0) flush();
}

template <int CH>
static cudaError_t launch_bwd_t(const ViewParams& vp, const uint2* ranges, const uint32_t* point_list,
                                const SplatRec* rec, const float* bg, const float* final_T,
                                const uint32_t* n_contrib, const float* dL_dpix, const float* dL_dfeat_pix,
                                const float* dL_ddepth, const BwdOut& out, cudaStream_t s) {
    const size_t smem = sizeof(BwdSmem);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(composite_bwd_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int chunks = CH > 0 ? (vp.C + CH - 1) / CH : 1;
    dim3 grid(vp.grid_x, vp.grid_y, chunks);
    int vec_io = 0;
    if (vp.W % 4 == 0 && (reinterpret_cast<uintptr_t>(dL_dfeat_pix) & 15) == 0) vec_io |= 1;
    if (vp.C % 4 == 0 && (reinterpret_cast<uintptr_t>(out.dL_dfeature) & 15) == 0) vec_io |= 2;
    composite_bwd_kernel<CH><<<grid, kBlockThreads, smem, s>>>(vp.W, vp.H, vp.C, ranges, point_list, rec, bg, final_T,
                                                              n_contrib, dL_dpix, dL_dfeat_pix, dL_ddepth, out,
                                                              vec_io);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_composite_bwd(const ViewParams& vp, const uint2* ranges, const uint32_t* point_list,
                                 const SplatRec* rec, const float* bg, const float* final_T,
                                 const uint32_t* n_contrib, const float* dL_dpix, const float* dL_dfeat_pix,
                                 const float* dL_ddepth, float* dL_dmean2D, float* dL_dconic,
                                 float* dL_dopacity, float* dL_dcolor, float* dL_dfeature, float* dL_dz,
                                 cudaStream_t s) {
    BwdOut out{dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dfeature, dL_dz};
    if (vp.C == 0)
        return launch_bwd_t<0>(vp, ranges, point_list, rec, bg, final_T, n_contrib, dL_dpix, dL_dfeat_pix, dL_ddepth,
                               out, s);
    if (vp.C <= 32)
        return launch_bwd_t<32>(vp, ranges, point_list, rec, bg, final_T, n_contrib, dL_dpix, dL_dfeat_pix, dL_ddepth,
                                out, s);
    if (vp.C <= 64)
        return launch_bwd_t<64>(vp, ranges, point_list, rec, bg, final_T, n_contrib, dL_dpix, dL_dfeat_pix, dL_ddepth,
                                out, s);
    return launch_bwd_t<128>(vp, ranges, point_list, rec, bg, final_T, n_contrib, dL_dpix, dL_dfeat_pix, dL_ddepth,
                             out, s);
}

}  // namespace f3dgs
