// tcgen05 kind::tf32 issue-rate probe (development tool): cycles per MMA for the operand configurations the composite
// kernels could use, one CTA on one SM, R back-to-back MMAs on (uninitialised) shared / tensor memory.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tc_rate tc_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)lt << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int amn, int bmn, int fmt) {  // fmt: 2 = tf32, 1 = bf16
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
struct P {
    long long* out;
    int mode, N, reps;
};
extern __shared__ __align__(1024) unsigned char smem_raw[];

__global__ void __launch_bounds__(128, 1) rate_kernel(P p) {
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 256; i += 128) reinterpret_cast<float*>(smem_raw)[i] = 1.0f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_s;
    if (tid == 0) {
        const uint32_t base = smem_u32(smem_raw);
        const int N = p.N;
        uint64_t a, b;
        uint32_t idesc;
        bool ts = false, bf16 = false;
        switch (p.mode) {
            case 0:  // SS, A and B MN-major SW128_BASE32B (forward TC kernel today)
                a = make_desc(base, 2048, 512, 1); b = make_desc(base + 65536, 2048, 512, 1); idesc = make_idesc(128, N, 1, 1, 2); break;
            case 1:  // SS, A and B K-major SW128
                a = make_desc(base, 16, 1024, 2); b = make_desc(base + 65536, 16, 1024, 2); idesc = make_idesc(128, N, 0, 0, 2); break;
            case 2:  // TS, B MN-major
                ts = true; b = make_desc(base + 65536, 2048, 512, 1); idesc = make_idesc(128, N, 0, 1, 2); break;
            case 3:  // TS, B K-major
                ts = true; b = make_desc(base + 65536, 16, 1024, 2); idesc = make_idesc(128, N, 0, 0, 2); break;
            case 4:  // SS, A K-major, B MN-major
                a = make_desc(base, 16, 1024, 2); b = make_desc(base + 65536, 2048, 512, 1); idesc = make_idesc(128, N, 0, 1, 2); break;
            case 5:  // bf16 SS K-major (K = 16)
                bf16 = true; a = make_desc(base, 16, 1024, 2); b = make_desc(base + 65536, 16, 1024, 2); idesc = make_idesc(128, N, 0, 0, 1); break;
            default:  // bf16 TS, B K-major
                bf16 = true; ts = true; b = make_desc(base + 65536, 16, 1024, 2); idesc = make_idesc(128, N, 0, 0, 1); break;
        }
        const long long t0 = clock64();
        for (int r = 0; r < p.reps; r++) {
            if (!ts) {
                if (!bf16)
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(1) : "memory");
                else
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(1) : "memory");
            } else {
                if (!bf16)
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, q;\n\t}" ::"r"(tmem), "r"(tmem + 256), "l"(b), "r"(idesc), "r"(1) : "memory");
                else
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, q;\n\t}" ::"r"(tmem), "r"(tmem + 256), "l"(b), "r"(idesc), "r"(1) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(smem_u32(&mbar)) : "memory");
        p.out[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long* d;
    cudaMalloc(&d, 8);
    const char* names[] = {"SS A,B MN-major (tf32)", "SS A,B K-major (tf32)", "TS, B MN-major (tf32)", "TS, B K-major (tf32)",
                           "SS A K-major, B MN-major (tf32)", "SS K-major (bf16, K=16)", "TS, B K-major (bf16, K=16)"};
    const int Ns[] = {256, 128, 64, 32, 16};
    for (int mode = 0; mode < 7; mode++)
        for (int N : Ns) {
            const int reps = 2000;
            P p{d, mode, N, reps};
            rate_kernel<<<1, 128, 200 * 1024>>>(p);
            cudaError_t e = cudaDeviceSynchronize();
            long long c = 0;
            cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            printf("%-34s M=128 N=%3d: %7.1f cycles per MMA  (%s)\n", names[mode], N, (double)c / reps, cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
