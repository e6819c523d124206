// Micro-benchmark: sustained FFMA issue rate on B200 for the register pattern of the composite feature loop
// (128 accumulators per thread, operands f[4] x w[4]), at 1, 2 and 4 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
template <int ITERS>
__global__ void __launch_bounds__(512, 1) k(float* out, const float4* in, int n) {
    float acc[32][4];
#pragma unroll
    for (int p = 0; p < 32; p++) for (int c = 0; c < 4; c++) acc[p][c] = 0.f;
    float4 f = in[threadIdx.x & 31];
    float4 w[8];
#pragma unroll
    for (int q = 0; q < 8; q++) w[q] = in[32 + q];
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const float wv[4] = {w[q].x, w[q].y, w[q].z, w[q].w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                acc[q * 4 + i][0] = fmaf(f.x, wv[i], acc[q * 4 + i][0]);
                acc[q * 4 + i][1] = fmaf(f.y, wv[i], acc[q * 4 + i][1]);
                acc[q * 4 + i][2] = fmaf(f.z, wv[i], acc[q * 4 + i][2]);
                acc[q * 4 + i][3] = fmaf(f.w, wv[i], acc[q * 4 + i][3]);
            }
        }
        f.x += 1e-9f;  // keep the loop from being hoisted
    }
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 32; p++) for (int c = 0; c < 4; c++) s += acc[p][c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out; float4* in;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&in, 64 * 16); cudaMemset(in, 0, 64 * 16);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int threads : {128, 256, 512}) {
        const int n = 20000;
        k<0><<<148, threads>>>(out, in, 10);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<0><<<148, threads>>>(out, in, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ffma_warp_instr = (double)n * 128 * (threads / 32) * 148;
        double per_smsp_per_us = ffma_warp_instr / (148 * 4) / (ms * 1e3);
        printf("threads/SM %d: %.3f ms, %.1f FFMA warp-instr per SMSP per us (clock %.0f MHz nominal => %.2f per cycle at max clock), %.1f TFLOP/s\n",
               threads, ms, per_smsp_per_us, clk / 1e3, per_smsp_per_us / (clk / 1e3), ffma_warp_instr * 64 / (ms * 1e-3) / 1e12);
    }
    return 0;
}
