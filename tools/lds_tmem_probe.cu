// tools/lds_tmem_probe.cu -- two hardware facts the element-thread remap of the tcgen05 kernel rests on (round 2):
//  (1) the shared-memory pipe cost of LDS.128 / LDS.64 by address pattern: warp-uniform (what "thread = instance" does
//      today), 4 distinct contiguous 16-byte records (lane & 3), 8 distinct, 32 distinct;
//  (2) which (TMEM lane, column) each register of tcgen05.ld.16x256b.x4 holds (the mma-style fragment: a thread gets two
//      instances x pairs of adjacent columns), checked by writing lane * 1000 + column with 32x32b stores first.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lds_tmem_probe tools/lds_tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VEC, int PATTERN>
__global__ void __launch_bounds__(1024) lds_kernel(int iters, long long* out, float* sink) {
    extern __shared__ __align__(16) unsigned char sm[];
    float* f = reinterpret_cast<float*>(sm);
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) f[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int sel;
    if (PATTERN == 0) sel = 0;                 // uniform
    else if (PATTERN == 1) sel = lane & 3;     // 4 distinct, contiguous records
    else if (PATTERN == 2) sel = lane & 7;     // 8 distinct
    else sel = lane;                           // all distinct
    const uint32_t base = smem_u32(sm) + sel * (VEC * 4);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t a = base + ((it * 16 + j) & 31) * 512;
            if (VEC == 4) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
                acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w;
            } else if (VEC == 2) {
                float2 v;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
                acc0 += v.x; acc1 += v.y;
            } else {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
                acc0 += v;
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc0 + acc1 + acc2 + acc3 == 1.2345f) sink[0] = acc0;
}

template <int VEC, int PATTERN>
static void run_lds(const char* name, int warps) {
    long long* d;
    float* s;
    cudaMalloc(&d, 8 * 148);
    cudaMalloc(&s, 4);
    const int iters = 2000;
    cudaFuncSetAttribute(lds_kernel<VEC, PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    lds_kernel<VEC, PATTERN><<<148, warps * 32, 65536>>>(iters, d, s);
    lds_kernel<VEC, PATTERN><<<148, warps * 32, 65536>>>(iters, d, s);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = 0;
    for (int i = 0; i < 148; ++i) cyc += (double)h[i];
    cyc /= 148;
    const double n = (double)iters * 16 * warps;  // warp-level LDS instructions per SM
    printf("LDS.%-3d %-22s warps=%2d  cycles per warp-instruction (SM-wide) %.3f\n", VEC * 32, name, warps, cyc / n);
    cudaFree(d);
    cudaFree(s);
}

// ---- TMEM layout -------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tmem_kernel(uint32_t* out) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tptr;
    const uint32_t tl = tbase + ((uint32_t)(warp * 32) << 16);
    // 32x32b store: thread = its own lane, 32 consecutive columns
    const int tm_lane = warp * 32 + lane;
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t v[8];
        for (int i = 0; i < 8; ++i) v[i] = tm_lane * 1000 + c0 + i;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tl + c0), "r"(v[0]), "r"(v[1]),
                     "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                     : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // 16x256b.x4: 16 lanes x 32 columns, at lane offsets 0 and 16 of this warp's quadrant
    for (int half = 0; half < 2; ++half) {
        uint32_t r[16];
        const uint32_t ta = tl + ((uint32_t)(half * 16) << 16) + 8;  // column offset 8: columns 8 .. 39
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[((warp * 2 + half) * 32 + lane) * 16 + i] = r[i];
    }
    // 16x256b store of (k, row, col) tags then 32x32b read-back: the store side of the same fragment
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(64u) : "memory");
}

int main() {
    printf("== LDS cost by address pattern (one CTA per SM) ==\n");
    for (int warps : {8, 16}) {
        if (warps == 8) {
            run_lds<4, 0>("uniform", 8); run_lds<4, 1>("4 contiguous records", 8); run_lds<4, 2>("8 contiguous records", 8); run_lds<4, 3>("32 distinct", 8);
            run_lds<2, 0>("uniform", 8); run_lds<2, 1>("4 contiguous records", 8); run_lds<2, 3>("32 distinct", 8);
            run_lds<1, 0>("uniform", 8); run_lds<1, 3>("32 distinct", 8);
        } else {
            run_lds<4, 0>("uniform", 16); run_lds<4, 1>("4 contiguous records", 16); run_lds<4, 3>("32 distinct", 16);
            run_lds<2, 0>("uniform", 16); run_lds<2, 1>("4 contiguous records", 16);
        }
    }
    printf("== tcgen05.ld.16x256b.x4 fragment (columns 8..39 requested) ==\n");
    uint32_t* d;
    cudaMalloc(&d, 4 * 2 * 32 * 16 * 4);
    tmem_kernel<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tmem kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    static uint32_t h[4 * 2 * 32 * 16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int warp = 0; warp < 4; ++warp)
        for (int half = 0; half < 2; ++half)
            for (int lane = 0; lane < 32; ++lane)
                for (int i = 0; i < 16; ++i) {
                    const uint32_t v = h[((warp * 2 + half) * 32 + lane) * 16 + i];
                    const int k = i >> 2, rsel = (i >> 1) & 1, e2 = i & 1;
                    const int row = warp * 32 + half * 16 + (lane >> 2) + 8 * rsel;
                    const int col = 8 + 8 * k + 2 * (lane & 3) + e2;
                    if (v != (uint32_t)(row * 1000 + col)) {
                        if (bad < 12) printf("  mismatch warp %d half %d lane %2d reg %2d: got lane %u col %u, expected lane %d col %d\n", warp, half, lane, i, v / 1000, v % 1000, row, col);
                        ++bad;
                    }
                }
    printf("16x256b.x4 layout hypothesis reg[4k + 2 rsel + e] = (lane base + lane/4 + 8 rsel, col base + 8 k + 2 (lane%%4) + e): %s (%d mismatches)\n",
           bad ? "WRONG" : "CONFIRMED", bad);
    printf("  lane 5 of warp 0, half 0:");
    for (int i = 0; i < 16; ++i) printf(" %u", h[(0 * 32 + 5) * 16 + i]);
    printf("\n");
    return 0;
}
