// tools/ffma2_probe.cu -- FP32 FMA issue rate: scalar FFMA vs packed fma.rn.f32x2 (FFMA2) on sm_100a
#include <cuda_runtime.h>
#include <cstdio>
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters) {
    float a[16];
    const float m = 1.0f + 1e-7f * threadIdx.x, c = 1e-9f;
    for (int i = 0; i < 16; ++i) a[i] = i + threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
    float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678f) out[0] = s;
}
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters) {
    unsigned long long a[8];
    const float m = 1.0f + 1e-7f * threadIdx.x, c = 1e-9f;
    unsigned long long mm, cc;
    asm("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    for (int i = 0; i < 8; ++i) { float x = i + threadIdx.x; asm("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(x)); }
#pragma unroll 1
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(mm), "l"(cc));
    float s = 0;
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i])); s += lo + hi; }
    if (s == 12345.678f) out[0] = s;
}
int main() {
    float* d; cudaMalloc(&d, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = 148 * 8;
    for (int v = 0; v < 2; ++v) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (v == 0) k_ffma<<<blocks, 256>>>(d, iters); else k_ffma2<<<blocks, 256>>>(d, iters);
            cudaEventRecord(e1); cudaDeviceSynchronize();
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 16 * 8 * (double)iters * 256 * blocks;
        printf("%s: %.3f ms  %.1f TFLOP/s  (%s)\n", v ? "fma.rn.f32x2" : "fma.rn.f32  ", ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
