// Throughput / latency of the legacy warp-level tensor-core product (mma.sync.m16n8k16 f16 -> f32, SASS HMMA.16816.F32) on
// sm_100a: cycles per instruction and per scheduler with U independent accumulator chains per warp and W warps per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hmma_probe tools/hmma_probe.cu && tools/hmma_probe
#include <cstdio>
#include <cuda_runtime.h>
template <int U>
__global__ void k(float* out, long long* cyc, int iters) {
    float d[U][4];
    for (int u = 0; u < U; ++u)
        for (int i = 0; i < 4; ++i) d[u][i] = 0.f;
    unsigned a0 = 0x3c003c00u + threadIdx.x, a1 = 0x3c003800u, a2 = 0x38003c00u, a3 = 0x3c003c00u, b0 = 0x3c003c00u, b1 = 0x38003800u;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < U; ++u)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[u][0]), "+f"(d[u][1]), "+f"(d[u][2]), "+f"(d[u][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int u = 0; u < U; ++u)
        for (int i = 0; i < 4; ++i) s += d[u][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int U>
void run(int warps) {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<U><<<148, 32 * warps>>>(out, cyc, iters);
    k<U><<<148, 32 * warps>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_warp = (double)h / (iters * U);
    const double per_sched = (double)h / ((double)iters * U * ((warps + 3) / 4));
    printf("warps/CTA %2d  chains/warp %d : %.1f cycles per HMMA per warp, %.2f cycles per HMMA per scheduler (%s)\n", warps, U, per_warp, per_sched,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 8, 16}) {
        run<1>(w);
        run<2>(w);
        run<4>(w);
        run<8>(w);
    }
    return 0;
}
