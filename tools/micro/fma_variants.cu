// Micro-benchmark: FP32 FMA issue rate on B200 for several operand patterns (all operands in vector registers).
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
__global__ void __launch_bounds__(256, 1) k(float* out, const float4* in, int n) {
    const int t = threadIdx.x;
    float4 f = in[t];
    float4 w[8];
#pragma unroll
    for (int q = 0; q < 8; q++) w[q] = in[256 + t * 8 + q];
    float s = 0.f;
    if (V == 1) {  // scalar FFMA, acc[32][4] += f[c] * w[p]
        float acc[32][4];
#pragma unroll
        for (int p = 0; p < 32; p++) for (int c = 0; c < 4; c++) acc[p][c] = 0.f;
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
            f.x += 1e-9f;
        }
#pragma unroll
        for (int p = 0; p < 32; p++) for (int c = 0; c < 4; c++) s += acc[p][c];
    } else if (V == 2) {  // scalar FFMA with only two distinct source registers: acc = fma(acc, w, w)
        float acc[32][4];
#pragma unroll
        for (int p = 0; p < 32; p++) for (int c = 0; c < 4; c++) acc[p][c] = 0.5f;
        for (int it = 0; it < n; it++) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float wv[4] = {w[q].x, w[q].y, w[q].z, w[q].w};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[q * 4 + i][c] = fmaf(acc[q * 4 + i][c], wv[i], wv[i]);
            }
        }
#pragma unroll
        for (int p = 0; p < 32; p++) for (int c = 0; c < 4; c++) s += acc[p][c];
    } else {  // packed FFMA2: accumulators are pixel pairs
        float2 acc[16][4];
#pragma unroll
        for (int p = 0; p < 16; p++) for (int c = 0; c < 4; c++) acc[p][c] = make_float2(0.f, 0.f);
        for (int it = 0; it < n; it++) {
            float2 fx, fy, fz, fw;
            if (V == 3) {  // explicit pairs
                fx = make_float2(f.x, f.x + 1.f); fy = make_float2(f.y, f.y + 1.f); fz = make_float2(f.z, f.z + 1.f); fw = make_float2(f.w, f.w + 1.f);
            } else {       // broadcast of one scalar
                fx = make_float2(f.x, f.x); fy = make_float2(f.y, f.y); fz = make_float2(f.z, f.z); fw = make_float2(f.w, f.w);
            }
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
#pragma unroll
        for (int p = 0; p < 16; p++) for (int c = 0; c < 4; c++) s += acc[p][c].x + acc[p][c].y;
    }
    out[blockIdx.x * blockDim.x + t] = s;
}
template <int V>
void run(float* out, const float4* in, const char* name) {
    const int n = 20000, threads = 256;
    k<V><<<148, threads>>>(out, in, 10);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<V><<<148, threads>>>(out, in, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)n * 128 * threads * 148;  // scalar FMAs
    printf("%-40s %.3f ms  %.1f TFLOP/s  (%.2f scalar-FMA lanes*32 per SMSP per ns)\n", name, ms, fma * 2 / (ms * 1e-3) / 1e12,
           fma / 32 / (148 * 4) / (ms * 1e6));
}
int main() {
    float *out; float4* in;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&in, (256 + 256 * 8) * 16); cudaMemset(in, 0, (256 + 256 * 8) * 16);
    run<1>(out, in, "V1 scalar FFMA f[c]*w[p]+acc");
    run<2>(out, in, "V2 scalar FFMA acc*w+w (2 src regs)");
    run<3>(out, in, "V3 FFMA2 explicit pairs");
    run<4>(out, in, "V4 FFMA2 broadcast x");
    return 0;
}
