// Shared machinery of the forward and backward composite kernels.
//
// One CTA renders one 16x16 tile (reference: one 256-thread block per tile, forward.cu:261-396).
// Here the CTA is 9 warps:
//   warps 0..7  consumers; warp w owns the 8x4 pixel block (w&1, w>>1) of the tile, one pixel
//               per lane in the alpha pass, one float4 of channels per lane in the feature pass
//   warp  8     producer: walks the tile's slice of the depth-sorted instance list, drops the
//               instances whose alpha>=1/255 footprint cannot reach the tile, and fills a ring of
//               shared-memory stages: the 48-byte splat records with 128-bit loads, the C-wide
//               feature rows with 1-D TMA bulk copies (cp.async.bulk -> SASS UBLKCP) that
//               complete on the stage's "full" mbarrier.
// Consumers release a stage through its "empty" mbarrier; no __syncthreads in the main loop, so
// the 8 pixel blocks drift apart freely (a block whose pixels are all saturated stops early).
#pragma once
#include "kernels.h"

namespace f3dgs {

constexpr int kConsumerWarps = 8;
constexpr int kBlockThreads = (kConsumerWarps + 1) * 32;
constexpr int kStageEntries = 32;

template <int CH>
struct alignas(128) Stage {
    float feat[kStageEntries][CH > 0 ? CH : 4];  // CH == 0: 512 B dummy, never touched
    float4 rec0[kStageEntries];                   // x, y, ex, ey
    float4 rec1[kStageEntries];                   // conic a, b, c, opacity
    float4 rec2[kStageEntries];                   // r, g, b, depth
    uint32_t listpos[kStageEntries];              // 1-based position in the tile's list (reference `contributor`)
    uint32_t gid[kStageEntries];                  // Gaussian index
    uint32_t n;                                   // valid entries
    uint32_t last;                                // 1 = no further stage follows
};

template <int CH, int STAGES>
struct alignas(128) Ring {
    Stage<CH> stage[STAGES];
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
    uint32_t done_mask;  // bit w set: consumer warp w needs no more instances
};

// pixel <-> lane mapping inside a warp's 8x4 block: 2x2 quads, quad q = lane>>2 laid out 4 across
__device__ __forceinline__ int lane_px(int lane) { return ((lane >> 2) & 3) * 2 + (lane & 1); }
__device__ __forceinline__ int lane_py(int lane) { return (lane >> 4) * 2 + ((lane >> 1) & 1); }

template <int CH, int STAGES>
__device__ __forceinline__ void ring_init(Ring<CH, STAGES>& ring) {
    // called by all threads before the role split; followed by __syncthreads()
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&ring.full[s], 1);
            mbar_init(&ring.empty[s], kConsumerWarps);
        }
        ring.done_mask = 0;
        mbar_fence_init();
    }
    if (CH > 0) {
        // rows shorter than CH (last channel chunk) rely on zero padding
        float4* p = reinterpret_cast<float4*>(&ring.stage[0]);
        const int n16 = (int)(sizeof(Stage<CH>) * STAGES / 16);
        for (int i = threadIdx.x; i < n16; i += blockDim.x) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        // order these generic-proxy stores before the async-proxy (bulk copy) writes to the same rows
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
}

// Producer warp.  REVERSE: walk the list back to front starting at `first` (backward pass).
// row_floats > 0: copy `row_floats` floats of features + gid*C + chunk_off per entry.
template <int CH, int STAGES, bool REVERSE>
__device__ __forceinline__ void producer_loop(Ring<CH, STAGES>& ring, const uint32_t* __restrict__ point_list,
                                              const SplatRec* __restrict__ rec, const float* __restrict__ features,
                                              int C, int chunk_off, int row_floats, bool use_bulk,
                                              uint32_t range_begin, uint32_t range_end, uint32_t walk_count,
                                              float tx0, float ty0, float tx1, float ty1) {
    const int lane = threadIdx.x & 31;
    int s = 0;
    uint32_t empty_parity = 1;  // fresh barrier: waiting on parity 1 falls through
    uint32_t fill = 0;
    const uint32_t row_bytes = (uint32_t)row_floats * 4u;

    mbar_wait(&ring.empty[0], empty_parity);

    auto list_index = [&](uint32_t i) -> uint32_t {  // i-th visited element -> index into point_list
        return REVERSE ? (range_begin + walk_count - 1 - i) : (range_begin + i);
    };
    auto publish = [&](uint32_t n, uint32_t last) {
        __syncwarp();
        if (lane == 0) {
            ring.stage[s].n = n;
            ring.stage[s].last = last;
            if (CH > 0 && use_bulk && n > 0)
                mbar_arrive_expect_tx(&ring.full[s], n * row_bytes);
            else
                mbar_arrive(&ring.full[s]);
        }
        __syncwarp();
    };
    auto advance = [&]() {
        s++;
        if (s == STAGES) {
            s = 0;
            empty_parity ^= 1;
        }
        mbar_wait(&ring.empty[s], empty_parity);
    };
    auto store_entry = [&](uint32_t slot, uint32_t gid, uint32_t lpos, float4 r0, float4 r1, float4 r2) {
        Stage<CH>& st = ring.stage[s];
        st.rec0[slot] = r0;
        st.rec1[slot] = r1;
        st.rec2[slot] = r2;
        st.listpos[slot] = lpos;
        st.gid[slot] = gid;
        if (CH > 0) {
            const float* src = features + (size_t)gid * C + chunk_off;
            if (use_bulk) {
                bulk_g2s(&st.feat[slot][0], src, row_bytes, &ring.full[s]);
            } else {
                for (int c = 0; c < row_floats; c++) st.feat[slot][c] = __ldg(src + c);
            }
        }
    };

    // two-deep software pipeline on the dependent loads (list index -> id -> record)
    const uint32_t nchunks = (walk_count + 31) / 32;
    uint32_t id_nxt = 0;     // ids of chunk c+1
    float4 a0, a1, a2;       // records of chunk c
    uint32_t id_cur = 0;
    a0 = a1 = a2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nchunks > 0) {
        if (lane < walk_count) id_cur = point_list[list_index(lane)];
        if (32 + lane < walk_count) id_nxt = point_list[list_index(32 + lane)];
        if (lane < walk_count) {
            const float4* r = reinterpret_cast<const float4*>(rec + id_cur);
            a0 = __ldg(r); a1 = __ldg(r + 1); a2 = __ldg(r + 2);
        }
    }
    for (uint32_t c = 0; c < nchunks; c++) {
        if (*reinterpret_cast<volatile uint32_t*>(&ring.done_mask) == (1u << kConsumerWarps) - 1u) break;
        // prefetch: records of chunk c+1, ids of chunk c+2
        float4 b0, b1, b2;
        b0 = b1 = b2 = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t i1 = (c + 1) * 32 + lane, i2 = (c + 2) * 32 + lane;
        uint32_t id_nn = 0;
        if (i1 < walk_count) {
            const float4* r = reinterpret_cast<const float4*>(rec + id_nxt);
            b0 = __ldg(r); b1 = __ldg(r + 1); b2 = __ldg(r + 2);
        }
        if (i2 < walk_count) id_nn = point_list[list_index(i2)];

        const uint32_t i0 = c * 32 + lane;
        const bool valid = i0 < walk_count;
        // does the alpha >= 1/255 footprint reach this tile?  (conservative, see alpha_extent)
        const bool keep = valid && (a0.x + a0.z >= tx0) && (a0.x - a0.z <= tx1) && (a0.y + a0.w >= ty0) &&
                          (a0.y - a0.w <= ty1);
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        const uint32_t cnt = __popc(m);
        const uint32_t rank = __popc(m & ((1u << lane) - 1u));
        const uint32_t lpos = list_index(i0) - range_begin + 1;
        const uint32_t room = kStageEntries - fill;
        if (keep && rank < room) store_entry(fill + rank, id_cur, lpos, a0, a1, a2);
        if (cnt >= room) {
            publish(kStageEntries, 0);
            advance();
            if (keep && rank >= room) store_entry(rank - room, id_cur, lpos, a0, a1, a2);
            fill = cnt - room;
        } else {
            fill += cnt;
        }
        a0 = b0; a1 = b1; a2 = b2;
        id_cur = id_nxt;
        id_nxt = id_nn;
    }
    publish(fill, 1);
}

}  // namespace f3dgs
