// Micro-benchmark: packed FP32 FMA (fma.rn.f32x2 -> FFMA2, sm_100) in the composite feature-loop pattern:
// accumulators are pixel pairs, x = (f_c, f_c), y = (w_2j, w_2j+1).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k(float* out, const float4* in, int n) {
    float2 acc[16][4];
#pragma unroll
    for (int p = 0; p < 16; p++) for (int c = 0; c < 4; c++) acc[p][c] = make_float2(0.f, 0.f);
    float4 f = in[threadIdx.x & 31];
    float4 w[8];
#pragma unroll
    for (int q = 0; q < 8; q++) w[q] = in[32 + q];
    for (int it = 0; it < n; it++) {
        const float2 fx = make_float2(f.x, f.x), fy = make_float2(f.y, f.y), fz = make_float2(f.z, f.z), fw = make_float2(f.w, f.w);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const float2 w01 = make_float2(w[q].x, w[q].y), w23 = make_float2(w[q].z, w[q].w);
            acc[2 * q][0] = __ffma2_rn(fx, w01, acc[2 * q][0]);
            acc[2 * q][1] = __ffma2_rn(fy, w01, acc[2 * q][1]);
            acc[2 * q][2] = __ffma2_rn(fz, w01, acc[2 * q][2]);
            acc[2 * q][3] = __ffma2_rn(fw, w01, acc[2 * q][3]);
            acc[2 * q + 1][0] = __ffma2_rn(fx, w23, acc[2 * q + 1][0]);
            acc[2 * q + 1][1] = __ffma2_rn(fy, w23, acc[2 * q + 1][1]);
            acc[2 * q + 1][2] = __ffma2_rn(fz, w23, acc[2 * q + 1][2]);
            acc[2 * q + 1][3] = __ffma2_rn(fw, w23, acc[2 * q + 1][3]);
        }
        f.x += 1e-9f;
    }
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 16; p++) for (int c = 0; c < 4; c++) s += acc[p][c].x + acc[p][c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out; float4* in;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&in, 64 * 16); cudaMemset(in, 0, 64 * 16);
    for (int threads : {128, 256, 512}) {
        const int n = 20000;
        k<<<148, threads>>>(out, in, 10);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<148, threads>>>(out, in, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double instr = (double)n * 64 * (threads / 32) * 148;
        printf("threads/SM %d: %.3f ms, %.1f FFMA2 warp-instr per SMSP per us, %.1f TFLOP/s\n", threads, ms,
               instr / (148 * 4) / (ms * 1e3), instr * 128 / (ms * 1e-3) / 1e12);
    }
    return 0;
}
