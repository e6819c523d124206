// tcgen05 kind::tf32 probe (development tool, not product code): validates, on the GPU box, the shared-memory
// operand layouts, descriptors and TMEM addressing that the tensor-core feature contraction relies on, and measures the
// accuracy of the error-compensated 3xTF32 product (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM).
//
//   case 0  "forward":      D[ch 128 x px 256]  = sum_inst F^T[ch,inst] * W[inst,px]
//                           A, B both MN-major, SWIZZLE_128B:  [block of 32][k][32 floats], chunk ^= (k & 7)
//   case 1  "backward SS":  D[ch 128 x inst 32] = sum_px dO^T[ch,px] * W[px,inst]
//                           A, B both K-major, SWIZZLE_128B:   [k-block of 32][row][32 floats], chunk ^= (row & 7)
//   case 2  "backward TS":  same product with A (dO^T) in tensor memory (tcgen05.st), B as in case 1
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;   // D format f32
    d |= 2u << 7;   // A format tf32
    d |= 2u << 10;  // B format tf32
    d |= (uint32_t)a_mn_major << 15;
    d |= (uint32_t)b_mn_major << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// physical float index inside a [rows][32 floats] SW128 region (1024-byte aligned): row r, element c (0..31)
__device__ __forceinline__ int sw128(int r, int c) { return r * 32 + ((((c >> 2) ^ (r & 7)) << 2) | (c & 3)); }

struct Params {
    const float* A;   // case 0: F[K][128] (row = instance);        case 1/2: dO[128][256] (row = channel)
    const float* B;   // case 0: W[K][256] (row = instance);        case 1/2: W[N=32][256] (row = instance)
    float* D;         // [128][N]
    int mode;         // 0, 1, 2
    int terms;        // 1: hi*hi only, 3: compensated
    int swap_lbo_sbo;
    int variant;      // mode 0 only: 0 = SW128 (16-byte chunks ^ (k & 7), atoms of 8 k);  1 = SW128_BASE32B (32-byte chunks ^ (k & 3), atoms of 4 k)
};
// float index of element c (0..31) of row r for the 32-byte-granular swizzle of MN-major 32-bit operands
__device__ __forceinline__ int sw128_32b(int r, int c) { return r * 32 + ((((c >> 3) ^ (r & 3)) << 3) | (c & 7)); }

constexpr int KF = 32;  // instances per stage in case 0
extern __shared__ __align__(1024) unsigned char smem_raw[];

__global__ void __launch_bounds__(128, 1) probe_kernel(Params p) {
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    float* sm = reinterpret_cast<float*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;

    if (p.mode == 0) {
        // smem: A_hi [4 blk][KF][32], A_lo, B_hi [8 blk][KF][32], B_lo   = (4+4+8+8)*KF*128 B = 96 KB
        float* Ahi = sm;
        float* Alo = Ahi + 4 * KF * 32;
        float* Bhi = Alo + 4 * KF * 32;
        float* Blo = Bhi + 8 * KF * 32;
        for (int i = tid; i < KF * 128; i += 128) {
            const int k = i / 128, c = i % 128;
            const float v = p.A[k * 128 + c], h = tf32_hi(v);
            const int o = (c >> 5) * (KF * 32) + (p.variant ? sw128_32b(k, c & 31) : sw128(k, c & 31));
            Ahi[o] = h;
            Alo[o] = v - h;
        }
        for (int i = tid; i < KF * 256; i += 128) {
            const int k = i / 256, c = i % 256;
            const float v = p.B[k * 256 + c], h = tf32_hi(v);
            const int o = (c >> 5) * (KF * 32) + (p.variant ? sw128_32b(k, c & 31) : sw128(k, c & 31));
            Bhi[o] = h;
            Blo[o] = v - h;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t idesc = make_idesc(128, 256, 1, 1);
            const uint32_t lt = p.variant ? 1u : 2u;
            uint32_t lbo = KF * 128, sbo = p.variant ? 512 : 1024;  // MN-major: LBO = next 32-element block along MN, SBO = next k atom
            if (p.swap_lbo_sbo) { uint32_t t = lbo; lbo = sbo; sbo = t; }
            uint32_t acc = 0;
            for (int g = 0; g < KF / 8; g++) {
                const uint64_t ah = make_desc(smem_u32(Ahi) + g * 1024, lbo, sbo, lt), al = make_desc(smem_u32(Alo) + g * 1024, lbo, sbo, lt);
                const uint64_t bh = make_desc(smem_u32(Bhi) + g * 1024, lbo, sbo, lt), bl = make_desc(smem_u32(Blo) + g * 1024, lbo, sbo, lt);
                mma_ss(tmem, ah, bh, idesc, acc);
                acc = 1;
                if (p.terms == 3) {
                    mma_ss(tmem, ah, bl, idesc, 1);
                    mma_ss(tmem, al, bh, idesc, 1);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
    } else {
        // smem: A_hi [8 kblk][128 ch][32] = 128 KB (mode 1: A_lo does not fit next to it -> mode 1 uses terms on B only + TS for lo)
        //       B_hi [8 kblk][32 inst][32], B_lo  = 2 * 32 KB
        constexpr int N = 32;
        float* Ahi = sm;
        float* Bhi = Ahi + 8 * 128 * 32;
        float* Blo = Bhi + 8 * N * 32;
        for (int i = tid; i < 128 * 256; i += 128) {
            const int m = i / 256, k = i % 256;
            const float v = p.A[m * 256 + k], h = tf32_hi(v);
            Ahi[(k >> 5) * (128 * 32) + sw128(m, k & 31)] = h;
        }
        for (int i = tid; i < N * 256; i += 128) {
            const int n = i / 256, k = i % 256;
            const float v = p.B[n * 256 + k], h = tf32_hi(v);
            const int o = (k >> 5) * (N * 32) + sw128(n, k & 31);
            Bhi[o] = h;
            Blo[o] = v - h;
        }
        // A_lo (mode 2: and A_hi too) into tensor memory: lane = channel, column = pixel; D sits at columns [0, 32)
        const uint32_t a_lo_col = 64, a_hi_col = 64 + 256 - 64;  // a_hi only used in mode 2 with K limited, see below
        (void)a_hi_col;
        {
            const int m = tid;  // 128 threads = 128 lanes; warp w owns lanes 32w..32w+31
            for (int c0 = 0; c0 < 256; c0 += 8) {
                uint32_t r[8];
                for (int j = 0; j < 8; j++) {
                    const float v = p.A[m * 256 + c0 + j];
                    r[j] = __float_as_uint(v - tf32_hi(v));
                }
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + a_lo_col + c0;
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
                             "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                             : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t idesc = make_idesc(128, N, 0, 0);
            uint32_t lbo = 16, sbo = 1024;  // K-major SW128: SBO = next 8 rows; LBO unused (1)
            if (p.swap_lbo_sbo) { uint32_t t = lbo; lbo = sbo; sbo = t; }
            uint32_t acc = 0;
            for (int j = 0; j < 8; j++)        // pixel block (one SW128 atom column of 32 K-elements)
                for (int t = 0; t < 4; t++) {  // 8 K-elements = 32 bytes inside the 128-byte swizzled row
                    const uint64_t ah = make_desc(smem_u32(Ahi) + j * (128 * 128) + t * 32, lbo, sbo);
                    const uint64_t bh = make_desc(smem_u32(Bhi) + j * (N * 128) + t * 32, lbo, sbo);
                    const uint64_t bl = make_desc(smem_u32(Blo) + j * (N * 128) + t * 32, lbo, sbo);
                    mma_ss(tmem, ah, bh, idesc, acc);
                    acc = 1;
                    if (p.terms == 3) {
                        mma_ss(tmem, ah, bl, idesc, 1);
                        if (p.mode == 2) mma_ts(tmem, tmem + a_lo_col + j * 32 + t * 8, bh, idesc, 1);
                    }
                }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
    }
    // everybody waits for the MMAs, then reads the accumulator: lane = row, column = n
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                : "=r"(ok)
                : "r"(smem_u32(&mbar))
                : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int N = p.mode == 0 ? 256 : 32;
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; j++) p.D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

int main() {
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int mode = 0; mode < 3; mode++)
        for (int terms = 1; terms <= 3; terms += 2)
            for (int swap = 0; swap < 2; swap++)
                for (int variant = 0; variant < (mode == 0 ? 2 : 1); variant++) {
                const int M = 128, N = mode == 0 ? 256 : 32, K = mode == 0 ? KF : 256;
                std::vector<float> A, B;
                srand(7 + mode);
                if (mode == 0) { A.resize(K * 128); B.resize(K * 256); } else { A.resize(128 * 256); B.resize(32 * 256); }
                for (auto& v : A) v = (float)frand();
                for (auto& v : B) v = (float)(0.5 * (frand() + 1.0));  // blend weights live in [0, 1)
                std::vector<double> ref((size_t)M * N, 0.0), mag((size_t)M * N, 0.0);
                for (int m = 0; m < M; m++)
                    for (int n = 0; n < N; n++) {
                        double s = 0, a = 0;
                        for (int k = 0; k < K; k++) {
                            const double x = mode == 0 ? A[k * 128 + m] : A[m * 256 + k];
                            const double y = mode == 0 ? B[k * 256 + n] : B[n * 256 + k];
                            s += x * y;
                            a += fabs(x * y);
                        }
                        ref[(size_t)m * N + n] = s;
                        mag[(size_t)m * N + n] = a;
                    }
                float *dA, *dB, *dD;
                cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, (size_t)M * N * 4);
                cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
                cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
                cudaMemset(dD, 0xFF, (size_t)M * N * 4);
                Params p{dA, dB, dD, mode, terms, swap, variant};
                probe_kernel<<<1, 128, 200 * 1024>>>(p);
                cudaError_t e = cudaDeviceSynchronize();
                std::vector<float> D((size_t)M * N);
                cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
                double worst = 0, worst_abs = 0;
                for (size_t i = 0; i < D.size(); i++) {
                    const double err = fabs((double)D[i] - ref[i]);
                    worst = fmax(worst, err / (mag[i] + 1e-30));
                    worst_abs = fmax(worst_abs, err);
                }
                printf("mode %d variant %d terms %d swap_lbo_sbo %d: cuda=%s  max |err| / sum|terms| = %.3e  max|err| = %.3e  D[0]=%g ref[0]=%g\n",
                       mode, variant, terms, swap, cudaGetErrorString(e), worst, worst_abs, D[0], ref[0]);
                cudaFree(dA); cudaFree(dB); cudaFree(dD);
                if (e != cudaSuccess) { printf("aborting after CUDA error\n"); return 1; }
            }
    return 0;
}
