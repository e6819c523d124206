// Micro-benchmark: throughput of red.global.add.{f32, v2.f32, v4.f32} in the pattern of the backward composite's
// feature-gradient scatter: every warp adds one contiguous row segment (32 lanes x VEC floats) to a pseudo-random row of a
// [rows, 128] fp32 matrix.  Reports payload GB/s and SM cycles per warp-level RED, for a matrix that fits in L2 (32 MB)
// and one that does not (512 MB, the config-3 dL_dfeature size).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

template <int VEC>
__global__ void __launch_bounds__(256) k(float* buf, unsigned rows_mask, int iters, int warps_rows) {
    const int lane = threadIdx.x & 31;
    unsigned x = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 2654435761u + 12345u;
    for (int it = 0; it < iters; it++) {
        x = x * 1664525u + 1013904223u;
        const unsigned row = (x >> 8) & rows_mask;
        float* p = buf + (size_t)row * 128 + lane * VEC;
        if (VEC == 4) asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
        if (VEC == 2) asm volatile("red.global.add.v2.f32 [%0], {%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
        if (VEC == 1) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
    }
}

template <int VEC>
void run(float* buf, unsigned rows, const char* what) {
    const int iters = 4000, grid = 148;
    k<VEC><<<grid, 256>>>(buf, rows - 1, 10, 0);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<VEC><<<grid, 256>>>(buf, rows - 1, iters, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double reds = (double)iters * 8 * grid;
    printf("%-28s VEC=%d rows=%8u : %.3f ms, %.1f GB/s payload, %.2f M warp-REDs/s/SM, %.1f ns per warp-RED per SM\n", what, VEC,
           rows, ms, reds * 32 * VEC * 4 / (ms * 1e-3) / 1e9, reds / grid / (ms * 1e-3) / 1e6, ms * 1e6 / (reds / grid));
}

int main() {
    float* buf;
    const size_t bytes = (size_t)1 << 29;  // 512 MB = 1M rows x 128 floats
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    run<4>(buf, 1u << 20, "v4, 512 MB (DRAM-backed)");
    run<4>(buf, 1u << 16, "v4, 32 MB (L2-resident)");
    run<2>(buf, 1u << 20, "v2, 512 MB");
    run<1>(buf, 1u << 20, "scalar, 512 MB");
    run<1>(buf, 1u << 16, "scalar, 32 MB");
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
